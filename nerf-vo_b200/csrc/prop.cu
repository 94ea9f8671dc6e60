// Fused proposal density field (HashMLPDensityField.density_fn, NS/fields/density_fields.py:93-116 through
// NS/fields/base_field.py:48-68): sample position -> SceneContraction -> (x+2)/4 -> selector -> L-level hash grid
// -> Linear(2L,16)+ReLU -> Linear(16,1) -> trunc_exp * selector, ONE kernel forward and ONE kernel backward.
//
// Why SIMT and not tcgen05: the network is 352 FLOP per sample next to 8*L scattered 8-byte gathers; the kernel is bound
// by L2 gather / atomic throughput, a 16-wide layer cannot fill a 128xNx16 UMMA tile, and keeping the whole chain in
// registers removes every intermediate tensor (positions, normalised x, selector, features, hidden activations) from HBM.
// The 64-wide field networks are the tensor-core path (mlp_tc.cu).
//
// Decomposition: one thread per sample; the 32 lanes of a warp are 32 consecutive samples of a ray, so gathers of the
// coarse levels hit the same sectors (L1) and the backward scatter collapses equal-row runs in the warp before issuing
// `red.global.add.v2.f32` (grid_common.cuh).  Weight gradients: every warp stages (dh, [f,1], dz, [h,1]) of its 32 samples in
// shared memory and the lanes reduce the outer products over the 32 samples into registers (lane = (hidden unit, half of
// the input columns)); registers are summed across the CTA's warps at the end and flushed with one atomicAdd per
// parameter per CTA.
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
    for (int a = 0; a < 3; ++a) p[a] = __fadd_rn(__ldg(o + 3 * r + a), __fdiv_rn(__fmul_rn(__ldg(d + 3 * r + a), se), 2.f));
}

// hidden layer + output pre-activation, accumulation order of the SIMT reference kernel (bias first, inputs ascending)
template <int L>
__device__ __forceinline__ float prop_mlp(const float* __restrict__ sp, const float* f, float* h) {
    using PL = PropLayout<L>;
    float z = sp[PL::B1];
#pragma unroll
    for (int j = 0; j < PH; ++j) {
        float acc = sp[PL::B0 + j];
#pragma unroll
        for (int i = 0; i < PL::IN; ++i) acc = fmaf(sp[PL::W0 + j * PL::IN + i], f[i], acc);
        h[j] = fmaxf(acc, 0.f);
    }
#pragma unroll
    for (int j = 0; j < PH; ++j) z = fmaf(sp[PL::W1 + j], h[j], z);
    return z;
}

template <int L, typename RowT>
__global__ void __launch_bounds__(256) k_prop_fwd(const __grid_constant__ GridP p, int64_t N, int S, const float* __restrict__ o, const float* __restrict__ d,
                                                  const float* __restrict__ starts, const float* __restrict__ ends, int64_t stride,
                                                  const float* __restrict__ positions, const RowT* __restrict__ table, const float* __restrict__ params,
                                                  float* __restrict__ density, float2* __restrict__ feat) {
    using PL = PropLayout<L>;
    __shared__ float sp[PL::NP];
    for (int e = threadIdx.x; e < PL::NP; e += blockDim.x) sp[e] = __ldg(params + e);
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    float pos[3], q[3];
    sample_point(t, S, o, d, starts, ends, stride, positions, pos);
    const float m = contract_point(pos, q);
    const uint32_t mask = (1u << p.log2T) - 1u;
    float f[PL::IN];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        const Corner c = make_corner(q[0], q[1], q[2], p.scale[l]);
        const float2 v = grid_level_forward(table + ((size_t)l << p.log2T), c, mask);
        f[2 * l] = v.x;
        f[2 * l + 1] = v.y;
        if (feat) feat[(int64_t)l * N + t] = v;  // level-major: a warp stores 256 contiguous bytes per level
    }
    float h[PH];
    const float z = prop_mlp<L>(sp, f, h);
    density[t] = expf(z) * m;  // trunc_exp forward (activations.py:33) * selector (density_fields.py:115)
}

#define PROP_BWD_THREADS 128
template <int L>
__global__ void __launch_bounds__(PROP_BWD_THREADS, 4) k_prop_bwd(const __grid_constant__ GridP p, int64_t N, int S, const float* __restrict__ o,
                                                               const float* __restrict__ d, const float* __restrict__ starts, const float* __restrict__ ends,
                                                               int64_t stride, const float* __restrict__ positions, const float* __restrict__ params,
                                                               const float2* __restrict__ feat, const float* __restrict__ ddensity,
                                                               float* __restrict__ dtable, float* __restrict__ dparams) {
    using PL = PropLayout<L>;
    constexpr int NW = PROP_BWD_THREADS / 32;
    __shared__ float sp[PL::NP];
    __shared__ float stage_all[NW * 32 * PL::RS];
    for (int e = threadIdx.x; e < PL::NP; e += blockDim.x) sp[e] = __ldg(params + e);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* stage = stage_all + warp * 32 * PL::RS;
    const uint32_t mask = (1u << p.log2T) - 1u;
    // wgrad roles: dW0ext[j][half*CH + c], c < CH  (column IN is the bias); lanes 0..PH: dW1ext[lane]
    const int wj = lane & (PH - 1), whalf = lane >> 4;
    float acc0[PL::CH];
#pragma unroll
    for (int c = 0; c < PL::CH; ++c) acc0[c] = 0.f;
    float acc1 = 0.f;

    for (int64_t base = ((int64_t)blockIdx.x * NW + warp) * 32; base < N; base += (int64_t)gridDim.x * PROP_BWD_THREADS) {
        const int64_t t = base + lane;
        const bool valid = t < N;
        const int64_t tt = valid ? t : N - 1;
        float pos[3], q[3];
        sample_point(tt, S, o, d, starts, ends, stride, positions, pos);
        const float m = contract_point(pos, q);
        float f[PL::IN];
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const float2 v = __ldg(feat + (int64_t)l * N + tt);
            f[2 * l] = v.x;
            f[2 * l + 1] = v.y;
        }
        float h[PH];
        const float z = prop_mlp<L>(sp, f, h);
        // d density / d z = selector * exp(clamp(z, -15, 15))   (activations.py:37-41)
        const float dz = valid ? __ldg(ddensity + tt) * m * expf(fminf(fmaxf(z, -15.f), 15.f)) : 0.f;
        if (__ballot_sync(0xffffffffu, dz != 0.f) == 0u) continue;  // nothing flows back from these 32 samples
        float dh[PH];
        float* row = stage + lane * PL::RS;
#pragma unroll
        for (int j = 0; j < PH; ++j) dh[j] = h[j] > 0.f ? dz * sp[PL::W1 + j] : 0.f;
#pragma unroll
        for (int i = 0; i < PL::IN; ++i) {  // dL/df, parked in this lane's staging row until the scatter
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < PH; ++j) a = fmaf(sp[PL::W0 + j * PL::IN + i], dh[j], a);
            row[PL::S_DF + i] = a;
        }
        // ---- weight gradients: stage this warp's 32 samples, reduce the outer products over them -----------------
        if (dparams) {
#pragma unroll
            for (int j = 0; j < PH; ++j) row[PL::S_DH + j] = dh[j];
#pragma unroll
            for (int i = 0; i < 2 * PL::CH; ++i) row[PL::S_F + i] = i < PL::IN ? f[i < PL::IN ? i : 0] : (i == PL::IN ? 1.f : 0.f);  // [f | 1 | 0-pad]
            row[PL::S_DZ] = dz;
#pragma unroll
            for (int j = 0; j < PH; ++j) row[PL::S_H + j] = h[j];
            row[PL::S_H + PH] = 1.f;
            __syncwarp();
#pragma unroll 4
            for (int s = 0; s < 32; ++s) {
                const float* r = stage + s * PL::RS;
                const float g = r[PL::S_DH + wj];
#pragma unroll
                for (int c = 0; c < PL::CH; ++c) acc0[c] = fmaf(g, r[PL::S_F + whalf * PL::CH + c], acc0[c]);
                if (lane <= PH) acc1 = fmaf(r[PL::S_DZ], r[PL::S_H + lane], acc1);
            }
            __syncwarp();
        }
        // ---- hash-table gradient: per level, warp-deduplicated scatter ------------------------------------------
        if (dtable) {
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                const Corner c = make_corner(q[0], q[1], q[2], p.scale[l]);
                grid_level_scatter(dtable + (((size_t)l << p.log2T) << 1), c, mask, row[PL::S_DF + 2 * l], row[PL::S_DF + 2 * l + 1], valid, lane);
            }
        }
    }
    if (!dparams) return;
    // ---- CTA reduction of the register accumulators, one atomicAdd per parameter -----------------------------------
    __syncthreads();
    float* red = stage_all;  // [NW][NP]
    float* mine = red + warp * PL::NP;
#pragma unroll
    for (int c = 0; c < PL::CH; ++c) {
        const int col = whalf * PL::CH + c;
        if (col < PL::IN)
            mine[PL::W0 + wj * PL::IN + col] = acc0[c];
        else if (col == PL::IN)
            mine[PL::B0 + wj] = acc0[c];
    }
    if (lane < PH)
        mine[PL::W1 + lane] = acc1;
    else if (lane == PH)
        mine[PL::B1] = acc1;
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

extern "C" int nvo_prop_density_forward(const nvo_grid_desc* g, int32_t hidden, void* stream, int64_t B, int32_t S, const float* origins,
                                        const float* directions, const float* starts, const float* ends, int64_t stride, const float* positions,
                                        const void* table, const float* params, float* density, float* feat) {
    GridP p;
    if (int e = prop_params(g, hidden, &p)) return e;
    NVO_CHECK(B >= 0 && S >= 1, "prop_density_forward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(table && params && density, "prop_density_forward: null pointer");
    NVO_CHECK(positions || (origins && directions && starts && ends), "prop_density_forward: need positions or rays + intervals");
    const int64_t N = B * S;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned int grid = nvo_blocks(N, 256);
    if (g->table_dtype == NVO_F32)
        k_prop_fwd<5, float2><<<grid, 256, 0, st>>>(p, N, S, origins, directions, starts, ends, stride, positions, (const float2*)table, params, density,
                                                   (float2*)feat);
    else
        k_prop_fwd<5, __half2><<<grid, 256, 0, st>>>(p, N, S, origins, directions, starts, ends, stride, positions, (const __half2*)table, params, density,
                                                    (float2*)feat);
    NVO_CUDA_LAUNCH_CHECK("prop_density_forward");
    return 0;
}

extern "C" int nvo_prop_density_backward(const nvo_grid_desc* g, int32_t hidden, void* stream, int64_t B, int32_t S, const float* origins,
                                         const float* directions, const float* starts, const float* ends, int64_t stride, const float* positions,
                                         const float* params, const float* feat, const float* ddensity, float* dtable, float* dparams) {
    GridP p;
    if (int e = prop_params(g, hidden, &p)) return e;
    NVO_CHECK(B >= 0 && S >= 1, "prop_density_backward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(params && feat && ddensity, "prop_density_backward: null pointer");
    NVO_CHECK(positions || (origins && directions && starts && ends), "prop_density_backward: need positions or rays + intervals");
    const int64_t N = B * S;
    const int64_t warps = (N + 31) / 32;
    const unsigned int grid = (unsigned int)min((warps + PROP_BWD_THREADS / 32 - 1) / (PROP_BWD_THREADS / 32), (int64_t)nvo_sm_count() * 12);
    k_prop_bwd<5><<<grid, PROP_BWD_THREADS, 0, (cudaStream_t)stream>>>(p, N, S, origins, directions, starts, ends, stride, positions, params,
                                                                      (const float2*)feat, ddensity, dtable, dparams);
    NVO_CUDA_LAUNCH_CHECK("prop_density_backward");
    return 0;
}
