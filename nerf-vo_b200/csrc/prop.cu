// Fused proposal density field (HashMLPDensityField.density_fn, NS/fields/density_fields.py:93-116 through
// NS/fields/base_field.py:48-68): sample position -> SceneContraction -> (x+2)/4 -> selector -> L-level hash grid
// -> Linear(2L,16)+ReLU -> Linear(16,1) -> trunc_exp * selector, ONE kernel forward and ONE kernel backward.
//
// Why SIMT and not tcgen05: the network is 352 FLOP per sample next to 8*L scattered 8-byte gathers; the kernel is bound
// by L2 gather / atomic throughput, a 16-wide layer cannot fill a 128xNx16 UMMA tile, and keeping the whole chain in
// registers removes every intermediate tensor (positions, normalised x, selector, features, hidden activations) from HBM.
// The 64-wide field networks are the tensor-core path (mlp_tc.cu).
//
// Decomposition: forward one thread per sample (the 32 lanes of a warp are 32 consecutive samples of a ray, so gathers of the
// coarse levels hit the same sectors in L1); backward one thread per 4 consecutive samples, which merges equal-cell runs in
// registers before issuing paired 16-byte reductions (ScatterRun, grid_common.cuh).  Weight gradients: every warp stages
// (dh, [f,1], dz, [h,1]) of 32 samples in shared memory and the lanes reduce the outer products over them into registers
// (lane = (hidden unit, half of the input columns)); registers are summed across the CTA's warps at the end and flushed with
// one atomicAdd per parameter per CTA.
#include "grid_common.cuh"

#define PH 16  // hidden width (nerfacto proposal networks, NS/models/nerfacto.py:93-97)

template <int L>
struct PropLayout {
    static constexpr int IN = 2 * L;
    static constexpr int W0 = 0;                 // [PH][IN] row-major (torch Linear.weight)
    static constexpr int B0 = PH * IN;           // [PH]
    static constexpr int W1 = B0 + PH;           // [1][PH]
    static constexpr int B1 = W1 + PH;           // [1]
    static constexpr int NP = B1 + 1;
    static constexpr int CH = (IN + 1 + 1) / 2;  // columns of [f,1] handled by one half-warp
    // staging row: dh[PH] | fext[2*CH] | dz | hext[PH+1] | df[IN]
    static constexpr int S_DH = 0, S_F = PH, S_DZ = PH + 2 * CH, S_H = S_DZ + 1, S_DF = S_H + PH + 1;
    static constexpr int RS = (S_DF + IN) | 1;  // odd stride: conflict-free row-per-lane writes
};

// position of sample (ray r, index k): rays.py:55 when `positions` is NULL, else the given point
__device__ __forceinline__ void sample_point(int64_t t, int S, const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ starts,
                                             const float* __restrict__ ends, int64_t stride, const float* __restrict__ positions, float* p) {
    if (positions) {
        p[0] = __ldg(positions + 3 * t);
        p[1] = __ldg(positions + 3 * t + 1);
        p[2] = __ldg(positions + 3 * t + 2);
        return;
    }
    const int64_t r = t / S;
    const int k = (int)(t - r * S);
    const float se = __fadd_rn(__ldg(starts + r * stride + k), __ldg(ends + r * stride + k));
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = __fadd_rn(__ldg(o + 3 * r + a), __fmul_rn(__fmul_rn(__ldg(d + 3 * r + a), se), 0.5f));
}

// MLP parameters live in constant memory (one slot per proposal network, refreshed from the fp32 parameter buffer by a
// device-to-device copy in front of every launch): every multiply-accumulate is then ONE FFMA with a uniform-register operand
// (LDCU.128 fetches four weights) instead of a shared-memory load + FFMA — these kernels are issue-bound, not bandwidth-bound.
// The backward keeps a SECOND copy (c_propb) for its dgrad: with one symbol ptxas sees every weight used twice (recompute + dgrad),
// hoists all 193 of them into vector registers and spills (168 registers, 300 B of spills before; 100 registers now).
#define PROP_SLOTS 4
#define PROP_SLOT_FLOATS 256
__constant__ float c_prop[PROP_SLOTS][PROP_SLOT_FLOATS];
__constant__ float c_propb[PROP_SLOTS][PROP_SLOT_FLOATS];

// Packed fp32 FMAs (sm_100 FFMA2: two IEEE fmas per instruction, each rounded like fmaf).  These kernels are issue-bound, not FMA-pipe-bound, and
// half of their instructions are the MLP's multiply-adds: pairing them frees issue slots for the hashing / address / interpolation work.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 cw2(const float* w) { return *reinterpret_cast<const float2*>(w); }  // two adjacent weights (8-byte aligned)

// hidden layer + output pre-activation.  Each output accumulates its even and its odd inputs in the two halves of one packed accumulator
// (bias in the even half, inputs ascending in each half), summed at the end; forward and the backward's recomputation share this function.
template <int L, int SLOT>
__device__ __forceinline__ float prop_mlp(const float* f, float* h) {
    using PL = PropLayout<L>;
    static_assert(PL::IN % 2 == 0 && PH % 2 == 0, "packed FMAs pair adjacent inputs / hidden units");
    float2 f2[PL::IN / 2];
#pragma unroll
    for (int i = 0; i < PL::IN / 2; ++i) f2[i] = make_float2(f[2 * i], f[2 * i + 1]);
#pragma unroll
    for (int j = 0; j < PH; ++j) {
        float2 acc = make_float2(c_prop[SLOT][PL::B0 + j], 0.f);
#pragma unroll
        for (int i = 0; i < PL::IN / 2; ++i) acc = ffma2(cw2(&c_prop[SLOT][PL::W0 + j * PL::IN + 2 * i]), f2[i], acc);
        h[j] = fmaxf(acc.x + acc.y, 0.f);
    }
    float2 z = make_float2(c_prop[SLOT][PL::B1], 0.f);
#pragma unroll
    for (int j = 0; j < PH / 2; ++j) z = ffma2(cw2(&c_prop[SLOT][PL::W1 + 2 * j]), make_float2(h[2 * j], h[2 * j + 1]), z);
    return z.x + z.y;
}

// Saved-for-backward layout (opaque `feat` buffer): features level-major [L][Npad] float2, then [Npad] float4 = (normalised position,
// selector).  Inside every block of 128 samples the sample j = 4*lane + g sits at g*32 + lane, so the backward (one thread = 4
// consecutive samples) reads 256 / 512 contiguous bytes per warp while the forward (one thread = one sample) still writes whole sectors.
__device__ __forceinline__ int64_t feat_slot(int64_t t) { return (t & ~(int64_t)127) + ((t & 3) << 5) + ((t & 127) >> 2); }

template <int L, int SLOT, typename RowT, bool PAIR>
__global__ void __launch_bounds__(256) k_prop_fwd(const __grid_constant__ GridP p, int64_t N, int64_t Npad, int S, const float* __restrict__ o,
                                                  const float* __restrict__ d, const float* __restrict__ starts, const float* __restrict__ ends,
                                                  int64_t stride, const float* __restrict__ positions, const RowT* __restrict__ table,
                                                  float* __restrict__ density, float2* __restrict__ feat) {
    using PL = PropLayout<L>;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    float pos[3], q[3];
    sample_point(t, S, o, d, starts, ends, stride, positions, pos);
    const float m = contract_point(pos, q);
    const uint32_t mask = (1u << p.log2T) - 1u;
    float f[PL::IN];
    const int64_t slot = feat_slot(t);
#pragma unroll
    for (int l = 0; l < L; ++l) {
        const Corner c = make_corner(q[0], q[1], q[2], p.scale[l]);
        const float2 v = grid_level_forward<RowT, PAIR>(table + ((size_t)l << p.log2T), c, mask);
        f[2 * l] = v.x;
        f[2 * l + 1] = v.y;
        if (feat) feat[(int64_t)l * Npad + slot] = v;
    }
    if (feat) reinterpret_cast<float4*>(feat + (int64_t)L * Npad)[slot] = make_float4(q[0], q[1], q[2], m);
    float h[PH];
    const float z = prop_mlp<L, SLOT>(f, h);
    density[t] = expf(z) * m;  // trunc_exp forward (activations.py:33) * selector (density_fields.py:115)
}

// ---------------------------------------------------------------------------------------------------------------------
// backward: one warp = 128 consecutive samples, one thread = G = 4 consecutive samples of a ray.
//   pass A (per g): recompute the MLP from the saved features, dz -> dh -> df; park df and the saved normalised position in shared
//                   memory.  Weight gradients: dW1 / db1 accumulate in the thread's own registers (17 FFMA per sample, reduced over
//                   the warp once at the end of the kernel); dW0 / db0 = sum_s dh[s] (x) [f[s], 1] goes through a 32-row staging
//                   tile: lane = (half-warp, pair of hidden units, half of the 12 columns) owns a 2 x 6 register tile and walks 16 of the
//                   32 staged rows (4 shared-memory loads per 12 FFMA).
//   pass B (per level): walk the thread's 4 samples in order with a CellRun (grid_common.cuh): partial sums stay in registers while
//                   the integer cell does not change (samples along a ray stay in a coarse cell for many steps), one 16-byte
//                   red.global.add.v4.f32 per (y,z) corner pair per run.
// ---------------------------------------------------------------------------------------------------------------------
#define PROP_BWD_THREADS 128
#define PROP_G 4

template <int L>
struct PropSmem {
    using PL = PropLayout<L>;
    static constexpr int RS = 28;                               // staging row: dh[16] | f[IN] 1 0-pad to 12 ; 112 B rows (16-byte aligned,
                                                                // 16-byte chunks of 8 consecutive rows fall in distinct banks)
    static constexpr int S_DH = 0, S_F = 16;
    static constexpr int DF = 0;                                // [G][IN][32]
    static constexpr int Q = DF + PROP_G * PL::IN * 32;         // [G][3][32]
    static constexpr int STAGE = Q + PROP_G * 3 * 32;           // [32][RS]
    static constexpr int PER_WARP = STAGE + 32 * RS;
    static_assert(PL::IN + 1 <= 12 && PH == 16, "staging tile is laid out for 16 hidden units and <= 11 extended input columns");
};

// SPLIT: the kernel stops after pass A and hands the feature gradients (fp32 tile-major [tile][2L][128], the layout of the tensor-core
// MLP's input gradient) and the normalised positions ([Npad,3]) to the long-run table scatter of grid.cu (k_grid_bwd_run: 16 consecutive
// samples of a ray per thread and level).  Measured (profiles/r01_ncu_step_kernels_s10.md): the fused kernel runs 16 warps per SM at
// 126 registers and issues ~2000 instructions per sample, most of them in pass B's divergent flushes; two lean kernels are faster.
template <int L, int SLOT, bool SPLIT>
__global__ void __launch_bounds__(PROP_BWD_THREADS, SPLIT ? 5 : 4) k_prop_bwd(const __grid_constant__ GridP p, int64_t N, int64_t Npad, int S,
                                                                  const float2* __restrict__ feat, const float* __restrict__ ddensity,
                                                                  float* __restrict__ dtable, float* __restrict__ dparams,
                                                                  float* __restrict__ df_tmf, float* __restrict__ xq) {
    using PL = PropLayout<L>;
    using SM = PropSmem<L>;
    constexpr int NW = PROP_BWD_THREADS / 32;
    extern __shared__ __align__(16) float smem_f[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ws = smem_f + warp * SM::PER_WARP;
    float* dfs = ws + SM::DF;
    float* qs = ws + SM::Q;
    float* stage = ws + SM::STAGE;
    const uint32_t mask = (1u << p.log2T) - 1u;
    const float4* qsaved = reinterpret_cast<const float4*>(feat + (int64_t)L * Npad);
    // dW0ext roles: half-warp hw walks rows {8a + 4hw + b}, hidden units 2jp, 2jp+1, columns 6ch .. 6ch+5 of [f | 1 | 0]
    const int hw = lane >> 4, jp = (lane & 15) >> 1, ch = lane & 1;
    float acc0[2][6];
#pragma unroll
    for (int c = 0; c < 6; ++c) acc0[0][c] = acc0[1][c] = 0.f;
    float accw1[PH], accb1 = 0.f;
#pragma unroll
    for (int j = 0; j < PH; ++j) accw1[j] = 0.f;
    const int64_t n_st = (N + 127) >> 7;
    for (int64_t st = (int64_t)blockIdx.x * NW + warp; st < n_st; st += (int64_t)gridDim.x * NW) {
        const int64_t base = st << 7;
        unsigned any = 0u;
        // ---------------- pass A ----------------
#pragma unroll 1
        for (int g = 0; g < PROP_G; ++g) {
            const int64_t t = base + lane * PROP_G + g;
            const bool valid = t < N;
            const int64_t sl = base + g * 32 + lane;  // == feat_slot(t)
            float4 q4 = __ldg(qsaved + sl);
            float f[PL::IN];
#pragma unroll
            for (int l = 0; l < L; ++l) {
                const float2 v = __ldg(feat + (int64_t)l * Npad + sl);
                f[2 * l] = valid ? v.x : 0.f;
                f[2 * l + 1] = valid ? v.y : 0.f;
            }
            if (!valid) q4 = make_float4(0.f, 0.f, 0.f, 0.f);  // rows past N were never written by the forward
            qs[(g * 3 + 0) * 32 + lane] = q4.x;
            qs[(g * 3 + 1) * 32 + lane] = q4.y;
            qs[(g * 3 + 2) * 32 + lane] = q4.z;
            float h[PH];
            const float z = prop_mlp<L, SLOT>(f, h);
            // d density / d z = selector * exp(clamp(z, -15, 15))   (activations.py:37-41)
            const float dz = valid ? __ldg(ddensity + t) * q4.w * expf(fminf(fmaxf(z, -15.f), 15.f)) : 0.f;
            const unsigned nz = __ballot_sync(0xffffffffu, dz != 0.f);
            any |= nz;
            float dh[PH];
#pragma unroll
            for (int j = 0; j < PH; ++j) dh[j] = h[j] > 0.f ? dz * c_propb[SLOT][PL::W1 + j] : 0.f;
            {
                float2 df[PL::IN / 2];  // df[i] = sum_j W0[j][i] dh[j], hidden units ascending: two adjacent inputs per packed FMA
#pragma unroll
                for (int i = 0; i < PL::IN / 2; ++i) df[i] = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < PH; ++j) {
                    const float2 d2 = make_float2(dh[j], dh[j]);
#pragma unroll
                    for (int i = 0; i < PL::IN / 2; ++i) df[i] = ffma2(cw2(&c_propb[SLOT][PL::W0 + j * PL::IN + 2 * i]), d2, df[i]);
                }
#pragma unroll
                for (int i = 0; i < PL::IN / 2; ++i) {
                    dfs[(g * PL::IN + 2 * i) * 32 + lane] = df[i].x;
                    dfs[(g * PL::IN + 2 * i + 1) * 32 + lane] = df[i].y;
                }
            }
            if (dparams && nz != 0u) {
                {
                    const float2 dz2 = make_float2(dz, dz);
#pragma unroll
                    for (int j = 0; j < PH / 2; ++j) {
                        const float2 a = ffma2(dz2, make_float2(h[2 * j], h[2 * j + 1]), make_float2(accw1[2 * j], accw1[2 * j + 1]));
                        accw1[2 * j] = a.x, accw1[2 * j + 1] = a.y;
                    }
                }
                accb1 += dz;
                float4* row = reinterpret_cast<float4*>(stage + lane * SM::RS);
#pragma unroll
                for (int k = 0; k < 4; ++k) row[k] = make_float4(dh[4 * k], dh[4 * k + 1], dh[4 * k + 2], dh[4 * k + 3]);
                row[4] = make_float4(f[0], f[1], f[2], f[3]);
                row[5] = make_float4(f[4], f[5], f[6], f[7]);
                row[6] = make_float4(f[8], f[9], 1.f, 0.f);
                __syncwarp();
#pragma unroll 4
                for (int k = 0; k < 16; ++k) {
                    const float* r = stage + ((k >> 2) * 8 + hw * 4 + (k & 3)) * SM::RS;
                    const float2 gj = *reinterpret_cast<const float2*>(r + SM::S_DH + 2 * jp);
                    const float2* fe = reinterpret_cast<const float2*>(r + SM::S_F + 6 * ch);
                    const float2 gx = make_float2(gj.x, gj.x), gy = make_float2(gj.y, gj.y);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float2 v = fe[c];
                        const float2 a = ffma2(gx, v, make_float2(acc0[0][2 * c], acc0[0][2 * c + 1]));
                        const float2 b = ffma2(gy, v, make_float2(acc0[1][2 * c], acc0[1][2 * c + 1]));
                        acc0[0][2 * c] = a.x, acc0[0][2 * c + 1] = a.y;
                        acc0[1][2 * c] = b.x, acc0[1][2 * c + 1] = b.y;
                    }
                }
                __syncwarp();
            }
        }
        if (SPLIT) {
            // coalesced hand-over: lane owns rows 4*lane .. 4*lane+3 of the tile = its four samples (g = 0..3)
            __syncwarp();
#pragma unroll
            for (int i = 0; i < PL::IN; ++i) {
                const float4 v = make_float4(dfs[(0 * PL::IN + i) * 32 + lane], dfs[(1 * PL::IN + i) * 32 + lane], dfs[(2 * PL::IN + i) * 32 + lane],
                                             dfs[(3 * PL::IN + i) * 32 + lane]);
                *reinterpret_cast<float4*>(df_tmf + (((st * PL::IN) + i) << 7) + 4 * lane) = v;
            }
            float4* xo = reinterpret_cast<float4*>(xq + 3 * (base + 4 * lane));
            xo[0] = make_float4(qs[0 * 32 + lane], qs[1 * 32 + lane], qs[2 * 32 + lane], qs[3 * 32 + lane]);
            xo[1] = make_float4(qs[4 * 32 + lane], qs[5 * 32 + lane], qs[6 * 32 + lane], qs[7 * 32 + lane]);
            xo[2] = make_float4(qs[8 * 32 + lane], qs[9 * 32 + lane], qs[10 * 32 + lane], qs[11 * 32 + lane]);
        }
        // ---------------- pass B ----------------
        if (!SPLIT && dtable && any != 0u) {
            __syncwarp();
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                float* slab = dtable + (((size_t)l << p.log2T) << 1);
                const float scale = p.scale[l];
                CellRun run;
                run.reset();
#pragma unroll
                for (int g = 0; g < PROP_G; ++g) {
                    const float g0 = dfs[(g * PL::IN + 2 * l) * 32 + lane], g1 = dfs[(g * PL::IN + 2 * l + 1) * 32 + lane];
                    if (g0 == 0.f && g1 == 0.f) continue;  // masked / invalid sample: contributes nothing, must not break a run either
                    run.add(slab, qs[(g * 3) * 32 + lane], qs[(g * 3 + 1) * 32 + lane], qs[(g * 3 + 2) * 32 + lane], scale, mask, g0, g1);
                }
                run.flush(slab);
            }
        }
        __syncwarp();
    }
    if (!dparams) return;
    // ---- warp reduction of the thread-local accumulators, CTA reduction through shared memory, one atomicAdd per parameter ----
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        acc0[0][c] += __shfl_xor_sync(0xffffffffu, acc0[0][c], 16);
        acc0[1][c] += __shfl_xor_sync(0xffffffffu, acc0[1][c], 16);
    }
#pragma unroll
    for (int j = 0; j < PH; ++j) accw1[j] = nvo_warp_sum(accw1[j]);
    accb1 = nvo_warp_sum(accb1);
    __syncthreads();
    float* red = smem_f;  // [NW][NP]
    float* mine = red + warp * PL::NP;
    if (hw == 0) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = 2 * jp + u;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const int col = 6 * ch + c;
                if (col < PL::IN)
                    mine[PL::W0 + j * PL::IN + col] = acc0[u][c];
                else if (col == PL::IN)
                    mine[PL::B0 + j] = acc0[u][c];
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < PH; ++j) mine[PL::W1 + j] = accw1[j];
        mine[PL::B1] = accb1;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < PL::NP; e += blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) v += red[w * PL::NP + e];
        if (v != 0.f) atomicAdd(dparams + e, v);
    }
}

// =====================================================================================================================
static int prop_params(const nvo_grid_desc* g, int32_t hidden, GridP* p) {
    NVO_CHECK(g != nullptr, "prop_density: null grid descriptor");
    NVO_CHECK(hidden == PH, "prop_density: fused kernel supports hidden width %d (got %d)", PH, hidden);
    NVO_CHECK(g->n_levels == 5, "prop_density: fused kernel supports 5 levels (got %d)", g->n_levels);
    NVO_CHECK(g->log2_T >= 1 && g->log2_T <= 30, "prop_density: log2_T=%d out of range [1,30]", g->log2_T);
    NVO_CHECK(g->table_dtype == NVO_F32 || g->table_dtype == NVO_F16, "prop_density: bad table_dtype %d", g->table_dtype);
    p->L = g->n_levels;
    p->log2T = g->log2_T;
    for (int i = 0; i < NVO_MAX_LEVELS; ++i) p->scale[i] = i < g->n_levels ? g->scalings[i] : 0.f;
    return 0;
}

extern "C" int nvo_prop_density_supported(int32_t n_levels, int32_t hidden, int32_t n_layers) { return n_levels == 5 && hidden == PH && n_layers == 2; }

extern "C" int64_t nvo_prop_density_feat_floats(int32_t n_levels, int64_t n) { return ((int64_t)n_levels * 2 + 4) * ((n + 127) / 128 * 128); }

// params == NULL: the caller has uploaded the network into `slot` with nvo_prop_density_upload and nothing else has used the slot since
static int upload_params(int slot, const float* params, cudaStream_t st, bool backward_copy = false) {
    NVO_CHECK(slot >= 0 && slot < PROP_SLOTS, "prop_density: slot %d out of range [0,%d)", slot, PROP_SLOTS);
    if (!params) return 0;
    cudaError_t e = cudaMemcpyToSymbolAsync(c_prop, params, sizeof(float) * PropLayout<5>::NP, sizeof(float) * PROP_SLOT_FLOATS * slot,
                                            cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess && backward_copy)
        e = cudaMemcpyToSymbolAsync(c_propb, params, sizeof(float) * PropLayout<5>::NP, sizeof(float) * PROP_SLOT_FLOATS * slot, cudaMemcpyDeviceToDevice, st);
    NVO_CHECK(e == cudaSuccess, "prop_density: parameter upload failed: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int nvo_prop_density_upload(void* stream, int32_t slot, const float* params) {
    NVO_CHECK(params, "prop_density_upload: null pointer");
    return upload_params(slot, params, (cudaStream_t)stream, true);
}

template <int SLOT>
static void launch_fwd(const nvo_grid_desc* g, const GridP& p, cudaStream_t st, int64_t N, int64_t Npad, int S, const float* o, const float* d, const float* s,
                       const float* e, int64_t stride, const float* positions, const void* table, float* density, float* feat) {
    const unsigned int grid = nvo_blocks(N, 256);
    static const int pair = nvo_env_int("NVO_PROP_FWD_PAIR", 0);  // 0 (default): eight plain gathers per level; 1: paired 16-byte loads behind a divergent branch (54 vs 41 us: the kernel is issue-bound)
    if (g->table_dtype == NVO_F32) {
        if (pair)
            k_prop_fwd<5, SLOT, float2, true><<<grid, 256, 0, st>>>(p, N, Npad, S, o, d, s, e, stride, positions, (const float2*)table, density, (float2*)feat);
        else
            k_prop_fwd<5, SLOT, float2, false><<<grid, 256, 0, st>>>(p, N, Npad, S, o, d, s, e, stride, positions, (const float2*)table, density, (float2*)feat);
    } else
        k_prop_fwd<5, SLOT, __half2, true><<<grid, 256, 0, st>>>(p, N, Npad, S, o, d, s, e, stride, positions, (const __half2*)table, density, (float2*)feat);
}

extern "C" int nvo_prop_density_forward(const nvo_grid_desc* g, int32_t hidden, int32_t slot, void* stream, int64_t B, int32_t S, const float* origins,
                                        const float* directions, const float* starts, const float* ends, int64_t stride, const float* positions,
                                        const void* table, const float* params, float* density, float* feat) {
    GridP p;
    if (int e = prop_params(g, hidden, &p)) return e;
    NVO_CHECK(B >= 0 && S >= 1, "prop_density_forward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(table && density, "prop_density_forward: null pointer");
    NVO_CHECK(positions || (origins && directions && starts && ends), "prop_density_forward: need positions or rays + intervals");
    const int64_t N = B * S, Npad = (N + 127) / 128 * 128;
    cudaStream_t st = (cudaStream_t)stream;
    if (int e = upload_params(slot, params, st)) return e;
    switch (slot) {
        case 0: launch_fwd<0>(g, p, st, N, Npad, S, origins, directions, starts, ends, stride, positions, table, density, feat); break;
        case 1: launch_fwd<1>(g, p, st, N, Npad, S, origins, directions, starts, ends, stride, positions, table, density, feat); break;
        case 2: launch_fwd<2>(g, p, st, N, Npad, S, origins, directions, starts, ends, stride, positions, table, density, feat); break;
        default: launch_fwd<3>(g, p, st, N, Npad, S, origins, directions, starts, ends, stride, positions, table, density, feat); break;
    }
    NVO_CUDA_LAUNCH_CHECK("prop_density_forward");
    return 0;
}

template <int SLOT>
static int launch_bwd(const GridP& p, cudaStream_t st, int64_t N, int64_t Npad, int S, const float* feat, const float* ddensity, float* dtable, float* dparams,
                      float* df_tmf, float* xq) {
    const size_t smem = sizeof(float) * PropSmem<5>::PER_WARP * (PROP_BWD_THREADS / 32);
    const int64_t blocks = ((N + 127) / 128 + PROP_BWD_THREADS / 32 - 1) / (PROP_BWD_THREADS / 32);
    if (df_tmf) {
        cudaError_t err = cudaFuncSetAttribute(k_prop_bwd<5, SLOT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        NVO_CHECK(err == cudaSuccess, "prop_density_backward: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
        static const int ctas_per_sm = nvo_env_int("NVO_PROP_BWD_CTAS", 10);  // grid cap in CTAs per SM (5 fit at once: 40 KB of shared memory each)
        const unsigned int grid = (unsigned int)min(blocks, (int64_t)nvo_sm_count() * ctas_per_sm);
        k_prop_bwd<5, SLOT, true><<<grid, PROP_BWD_THREADS, smem, st>>>(p, N, Npad, S, (const float2*)feat, ddensity, nullptr, dparams, df_tmf, xq);
        return 0;
    }
    cudaError_t err = cudaFuncSetAttribute(k_prop_bwd<5, SLOT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    NVO_CHECK(err == cudaSuccess, "prop_density_backward: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    const unsigned int grid = (unsigned int)min(blocks, (int64_t)nvo_sm_count() * 8);
    k_prop_bwd<5, SLOT, false><<<grid, PROP_BWD_THREADS, smem, st>>>(p, N, Npad, S, (const float2*)feat, ddensity, dtable, dparams, nullptr, nullptr);
    return 0;
}

static int prop_backward_impl(const nvo_grid_desc* g, int32_t hidden, int32_t slot, void* stream, int64_t B, int32_t S, const float* origins,
                              const float* directions, const float* starts, const float* ends, int64_t stride, const float* positions,
                              const float* params, const float* feat, const float* ddensity, float* dtable, float* dparams, float* df_tmf, float* xq);

extern "C" int nvo_prop_density_backward(const nvo_grid_desc* g, int32_t hidden, int32_t slot, void* stream, int64_t B, int32_t S, const float* origins,
                                         const float* directions, const float* starts, const float* ends, int64_t stride, const float* positions,
                                         const float* params, const float* feat, const float* ddensity, float* dtable, float* dparams) {
    return prop_backward_impl(g, hidden, slot, stream, B, S, origins, directions, starts, ends, stride, positions, params, feat, ddensity, dtable, dparams,
                              nullptr, nullptr);
}

extern "C" int nvo_prop_density_backward_split(const nvo_grid_desc* g, int32_t hidden, int32_t slot, void* stream, int64_t B, int32_t S,
                                               const float* params, const float* feat, const float* ddensity, float* dparams, float* dfeat_tmf,
                                               float* xq) {
    NVO_CHECK(dfeat_tmf && xq, "prop_density_backward_split: null output pointer");
    // any non-null `positions` stands for "no ray arguments needed": the backward works from the saved features only
    return prop_backward_impl(g, hidden, slot, stream, B, S, nullptr, nullptr, nullptr, nullptr, 0, feat, params, feat, ddensity, nullptr, dparams, dfeat_tmf,
                              xq);
}

static int prop_backward_impl(const nvo_grid_desc* g, int32_t hidden, int32_t slot, void* stream, int64_t B, int32_t S, const float* origins,
                              const float* directions, const float* starts, const float* ends, int64_t stride, const float* positions,
                              const float* params, const float* feat, const float* ddensity, float* dtable, float* dparams, float* df_tmf, float* xq) {
    GridP p;
    if (int e = prop_params(g, hidden, &p)) return e;
    NVO_CHECK(B >= 0 && S >= 1, "prop_density_backward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(feat && ddensity, "prop_density_backward: null pointer");
    NVO_CHECK(positions || (origins && directions && starts && ends), "prop_density_backward: need positions or rays + intervals");
    const int64_t N = B * S, Npad = (N + 127) / 128 * 128;
    cudaStream_t st = (cudaStream_t)stream;
    if (int e = upload_params(slot, params, st, true)) return e;  // re-uploaded: another network may have used the slot since the forward
    int rc;
    switch (slot) {
        case 0: rc = launch_bwd<0>(p, st, N, Npad, S, feat, ddensity, dtable, dparams, df_tmf, xq); break;
        case 1: rc = launch_bwd<1>(p, st, N, Npad, S, feat, ddensity, dtable, dparams, df_tmf, xq); break;
        case 2: rc = launch_bwd<2>(p, st, N, Npad, S, feat, ddensity, dtable, dparams, df_tmf, xq); break;
        default: rc = launch_bwd<3>(p, st, N, Npad, S, feat, ddensity, dtable, dparams, df_tmf, xq); break;
    }
    if (rc) return rc;
    NVO_CUDA_LAUNCH_CHECK("prop_density_backward");
    return 0;
}
