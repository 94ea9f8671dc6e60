// Fully fused MLP on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM) for the 64-wide field networks.
//
// Data layout ("TMH", tile-major half): an activation matrix [n, K] lives in HBM as tiles of 128 rows, each tile a
// contiguous K*256-byte block [chunk = col/8][row][8 halfs] — exactly the UMMA canonical no-swizzle operand layout.  A tile
// therefore moves HBM -> shared memory with ONE bulk async copy (cp.async.bulk, the TMA engine's 1-D mode) that signals an
// mbarrier, producers (hash-grid gather, input assembly) write it with fully coalesced 16-byte stores, and neither
// transposed weights nor transposed activations are ever materialised:
//   * as a K-major operand (forward: A = activations, B = W[out][in]):       LBO = chunk stride, SBO = 128 B
//   * as an MN-major operand (dgrad: B = the SAME W tile; wgrad: A = dZ, B = activations, K = the 128 samples):
//                                                                             LBO = 128 B, SBO = chunk stride
// (descriptor conventions pinned on B200 by tools/umma_probe.cu).
//
// Weights: packed once per parameter update (k_tc_pack) into an fp16 image in the same tile layout + fp32 biases; every CTA
// pulls the image into shared memory with one bulk copy.
//
// Forward (k_mlp_tc_fwd): persistent CTAs of 128 threads (thread == sample row), 3-5 CTAs per SM.  The next tile's input is
// prefetched into the second input buffer while the current tile runs its layer chain: per layer ONE thread issues K/16
// tcgen05.mma (M = 128 samples, N = layer width) and commits to an mbarrier; the 4 warps pull the accumulator out of TMEM
// (tcgen05.ld), add bias, activate, convert to fp16 and write the row back to shared memory as the next layer's A operand
// (and to the saved-activation buffer, TMH layout, for the backward).
//
// Backward (k_mlp_tc_bwd): one CTA per SM (the weight-gradient accumulators of ALL layers stay resident in TMEM across every
// tile of the CTA: dW_l += dZ_l^T A_{l-1}, bias gradient through an appended ones-column, flushed once at the end).  All
// activations a tile needs (input + every saved hidden layer) are prefetched for the NEXT tile by bulk copies into the
// second stage while the current tile computes.  Per layer the dgrad MMAs are issued first and committed; the wgrad MMAs
// are issued right behind them and run on the tensor pipe while the 4 warps do the dgrad epilogue (TMEM -> x act' -> fp16
// -> the other dZ buffer), so the tensor pipe and the epilogue overlap.  Gradients are scaled by a device-side power of
// two (2^8 / max|dy|) before the fp16 conversion and unscaled in fp32 on the way out.
#include <cuda_fp16.h>
#include "nvo_common.cuh"

#include "tc_common.cuh"
#define NT 256           // threads per CTA: warp w owns TMEM lanes 32*(w&3).. (rows of the tile) and the column half (w>>2)
#define MAXW 64
#define TC_MAX_LAYERS 4
#define DW_COLS 80       // TMEM columns reserved per layer for dW (Kpad + 16 <= 80)

struct TcP {
    int n_layers, in_dim, k0pad;
    int dims[TC_MAX_LAYERS], npad[TC_MAX_LAYERS], kpad[TC_MAX_LAYERS], acts[TC_MAX_LAYERS];
    int w_off[TC_MAX_LAYERS], b_off[TC_MAX_LAYERS];      // float offsets into params (torch layout)
    int iw_off[TC_MAX_LAYERS], ib_off[TC_MAX_LAYERS];    // byte offsets into the packed weight image
    int img_bytes;
    int saved_chunk_off[TC_MAX_LAYERS];                  // chunk offset of layer l's output inside a saved tile
    int saved_chunks;                                    // chunks per saved tile
    int64_t x_tile_bytes, saved_tile_bytes;              // backward: byte strides between consecutive tiles of the input / saved buffers
    int a_off[TC_MAX_LAYERS];                            // backward: byte offset of layer l's INPUT tile inside the stage
    int stage_bytes;                                     // backward: bytes of the activation stage (every layer's input tile + ones/zero chunks)
    int wg_t[TC_MAX_LAYERS];                             // backward: 1 = weight gradient accumulated transposed (D[in][out], N = 16 columns)
    int dw_col[TC_MAX_LAYERS];                           // backward: first TMEM column of layer l's weight-gradient accumulator
    int tmem_cols;                                       // backward: TMEM columns to allocate (256 or 512)
    int n_params;
};

__device__ __forceinline__ float tc_act_fwd(float v, int act) {
    switch (act) {
        case NVO_ACT_RELU: return fmaxf(v, 0.f);
        case NVO_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case NVO_ACT_TANH: return tanhf(v);
        case NVO_ACT_EXP:
        case NVO_ACT_TRUNC_EXP: return expf(v);
        default: return v;
    }
}
__device__ __forceinline__ float tc_act_bwd(float a, int act) {
    switch (act) {
        case NVO_ACT_RELU: return a > 0.f ? 1.f : 0.f;
        case NVO_ACT_SIGMOID: return a * (1.f - a);
        case NVO_ACT_TANH: return 1.f - a * a;
        case NVO_ACT_EXP: return a;
        case NVO_ACT_TRUNC_EXP: return fminf(fmaxf(a, 3.0590232050182579e-07f), 3269017.3724721107f);
        default: return 1.f;
    }
}

__device__ __forceinline__ uint4 pack8(const float* v) {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]), c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    u.z = *reinterpret_cast<uint32_t*>(&c);
    u.w = *reinterpret_cast<uint32_t*>(&d);
    return u;
}


// ---- epilogue building blocks, specialised on the activation so the per-element code is branch-free ---------------
template <int ACT>
__device__ __forceinline__ float act_fwd_t(float v) {
    if (ACT == NVO_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == NVO_ACT_NONE) return v;
    return tc_act_fwd(v, ACT);
}
template <int ACT>
__device__ __forceinline__ float act_bwd_t(float a) {
    if (ACT == NVO_ACT_RELU) return a > 0.f ? 1.f : 0.f;
    if (ACT == NVO_ACT_NONE) return 1.f;
    return tc_act_bwd(a, ACT);
}

// column range of a K- or N-wide accumulator that column-half `half` (warps 0-3 / 4-7) of the CTA handles
__device__ __forceinline__ void half_range(int width, int half, int& c0, int& c1) {
    const int split = width > 32 ? 32 : (width == 32 ? 16 : width);
    c0 = half ? split : 0;
    c1 = half ? width : split;
}

// hidden layer of the forward: NC (16 | 32) accumulator columns -> + bias -> activation -> fp16 -> next layer's A tile (+ saved)
template <int ACT, int NC>
__device__ __forceinline__ void fwd_hidden_cols(uint32_t taddr, const float* __restrict__ sb, unsigned char* __restrict__ sAct, uint4* __restrict__ saved_tile,
                                                int c8_0, int r) {
    float v[NC];
    if (NC == 32)
        tmem_ld32(taddr, v);
    else
        tmem_ld16(taddr, v);
#pragma unroll
    for (int q = 0; q < NC / 4; ++q) {
        const float4 b = *reinterpret_cast<const float4*>(sb + 4 * q);  // warp-broadcast
        v[4 * q] = act_fwd_t<ACT>(v[4 * q] + b.x);
        v[4 * q + 1] = act_fwd_t<ACT>(v[4 * q + 1] + b.y);
        v[4 * q + 2] = act_fwd_t<ACT>(v[4 * q + 2] + b.z);
        v[4 * q + 3] = act_fwd_t<ACT>(v[4 * q + 3] + b.w);
    }
#pragma unroll
    for (int q = 0; q < NC / 8; ++q) {
        const uint4 u = pack8(v + q * 8);
        *reinterpret_cast<uint4*>(sAct + (c8_0 + q) * CHUNK_B + r * 16) = u;
        if (saved_tile) saved_tile[(c8_0 + q) * TM + r] = u;  // TMH: coalesced 16-byte stores
    }
}
template <int ACT>
__device__ __forceinline__ void fwd_hidden_layer(uint32_t trow, int np, int half, const float* __restrict__ sb, unsigned char* __restrict__ sAct,
                                                 uint4* __restrict__ saved_tile, int r) {
    int c, c1;
    half_range(np, half, c, c1);
    for (; c + 32 <= c1; c += 32) fwd_hidden_cols<ACT, 32>(trow + c, sb + c, sAct, saved_tile, c >> 3, r);
    if (c < c1) fwd_hidden_cols<ACT, 16>(trow + c, sb + c, sAct, saved_tile, c >> 3, r);
}

// hidden layer of the backward: NC dgrad columns -> x act'(a) -> fp16 -> the next dZ tile
template <int ACT, int NC>
__device__ __forceinline__ void bwd_hidden_cols(uint32_t taddr, const unsigned char* __restrict__ sA, unsigned char* __restrict__ sGn, int c8_0, int r) {
    float v[NC];
    if (NC == 32)
        tmem_ld32(taddr, v);
    else
        tmem_ld16(taddr, v);
#pragma unroll
    for (int q = 0; q < NC / 8; ++q) {
        if (ACT != NVO_ACT_NONE) {
            const uint4 a = *reinterpret_cast<const uint4*>(sA + (c8_0 + q) * CHUNK_B + r * 16);
            const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                v[q * 8 + 2 * j] *= act_bwd_t<ACT>(f.x);
                v[q * 8 + 2 * j + 1] *= act_bwd_t<ACT>(f.y);
            }
        }
        *reinterpret_cast<uint4*>(sGn + (c8_0 + q) * CHUNK_B + r * 16) = pack8(v + q * 8);
    }
}
template <int ACT>
__device__ __forceinline__ void bwd_hidden_layer(uint32_t trow, int kp, int half, const unsigned char* __restrict__ sA, unsigned char* __restrict__ sGn, int r) {
    int c, c1;
    half_range(kp, half, c, c1);
    for (; c + 32 <= c1; c += 32) bwd_hidden_cols<ACT, 32>(trow + c, sA, sGn, c >> 3, r);
    if (c < c1) bwd_hidden_cols<ACT, 16>(trow + c, sA, sGn, c >> 3, r);
}

// ================================================================================================================
// weight image: per layer W_l as an fp16 natural tile with npad rows (zero padded) followed by the fp32 bias [npad]
// ================================================================================================================
__global__ void __launch_bounds__(256) k_tc_pack(const __grid_constant__ TcP p, const float* __restrict__ params, unsigned char* __restrict__ img) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int l = 0; l < p.n_layers; ++l) {
        const int K = l == 0 ? p.in_dim : p.dims[l - 1], N = p.dims[l], kp = p.kpad[l], np = p.npad[l];
        __half* iw = reinterpret_cast<__half*>(img + p.iw_off[l]);
        float* ib = reinterpret_cast<float*>(img + p.ib_off[l]);
        for (int e = t; e < np * kp; e += stride) {
            const int o = e / kp, i = e - o * kp;
            const float w = (o < N && i < K) ? __ldg(params + p.w_off[l] + o * K + i) : 0.f;
            iw[((i >> 3) * np + o) * 8 + (i & 7)] = __float2half_rn(w);
        }
        for (int o = t; o < np; o += stride) ib[o] = o < N ? __ldg(params + p.b_off[l] + o) : 0.f;
    }
}

// ================================================================================================================
// forward.  smem: [in0 16 KB][in1 16 KB][act 16 KB][weight image][mbar_mma, mbar_w, mbar_in0, mbar_in1, tmem ptr]
// 256 threads: warp w reads TMEM lanes 32*(w&3).. (tile rows) and handles the column half (w>>2) of every epilogue, so a
// layer's TMEM -> bias -> activation -> fp16 -> shared chain is half as long per warp and twice as many warps hide it.
// ================================================================================================================
#define FWD_IN0 0
#define FWD_IN1 (8 * CHUNK_B)
#define FWD_ACT (16 * CHUNK_B)
#define FWD_W (24 * CHUNK_B)

__global__ void __launch_bounds__(NT) k_mlp_tc_fwd(const __grid_constant__ TcP p, int64_t n, const unsigned char* __restrict__ x16,
                                                   const unsigned char* __restrict__ wimg, const float* __restrict__ row_mask, float* __restrict__ y,
                                                   uint4* __restrict__ saved, int smem_ctrl_off) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sAct = smem + FWD_ACT;
    unsigned char* sW = smem + FWD_W;
    uint64_t* mbar_mma = reinterpret_cast<uint64_t*>(smem + smem_ctrl_off);
    uint64_t* mbar_w = mbar_mma + 1;
    uint64_t* mbar_in = mbar_mma + 2;  // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(mbar_mma + 4);
    const int tid = threadIdx.x, warp = tid >> 5, half = warp >> 2;
    const int r = ((warp & 3) << 5) | (tid & 31);  // row of the tile == TMEM lane
    const int64_t n_tiles = (n + TM - 1) / TM;
    const uint32_t in_bytes = (uint32_t)p.k0pad * 256u;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(mbar_mma, 1);
        mbar_init(mbar_w, 1);
        mbar_init(mbar_in, 1);
        mbar_init(mbar_in + 1, 1);
        fence_mbar_init();
        fence_async_smem();
        mbar_expect_tx(mbar_w, (uint32_t)p.img_bytes);
        bulk_g2s(sW, wimg, (uint32_t)p.img_bytes, mbar_w);
        mbar_expect_tx(mbar_in, in_bytes);
        bulk_g2s(smem + FWD_IN0, x16 + (int64_t)blockIdx.x * in_bytes, in_bytes, mbar_in);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    mbar_wait(mbar_w, 0);
    uint32_t phase = 0;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const int64_t row = tile * TM + r;
        const bool live = row < n;
        unsigned char* sIn = smem + (buf ? FWD_IN1 : FWD_IN0);
        // prefetch the next tile's input: the other buffer was last read by layer-0 MMAs of the previous tile (completed)
        if (tid == 0 && tile + gridDim.x < n_tiles) {
            mbar_expect_tx(mbar_in + (buf ^ 1), in_bytes);
            bulk_g2s(smem + (buf ? FWD_IN0 : FWD_IN1), x16 + (tile + gridDim.x) * in_bytes, in_bytes, mbar_in + (buf ^ 1));
        }
        mbar_wait(mbar_in + buf, (uint32_t)((it >> 1) & 1));
        for (int l = 0; l < p.n_layers; ++l) {
            const int kp = p.kpad[l], np = p.npad[l];
            if (warp == 0 && elect_one()) {  // elect.sync, not `tid == 0`: see elect_one() in tc_common.cuh
                tc_fence_after();
                const uint32_t idesc = umma_idesc(TM, np, 0, 0);
                const uint32_t a0 = smem_u32(l == 0 ? sIn : sAct), b0 = smem_u32(sW + p.iw_off[l]);
                for (int k = 0; k < (kp >> 4); ++k)
                    umma_f16(tmem, umma_desc(a0 + k * 2 * CHUNK_B, CHUNK_B, 128), umma_desc(b0 + k * 2 * np * 16, np * 16, 128), idesc, k > 0);
                umma_commit(mbar_mma);
            }
            mbar_wait(mbar_mma, phase);
            phase ^= 1;
            tc_fence_after();
            const float* sb = reinterpret_cast<const float*>(sW + p.ib_off[l]);
            const bool last = l == p.n_layers - 1;
            const int act = p.acts[l];
            if (!last) {
                uint4* saved_tile = saved ? saved + (tile * p.saved_chunks + p.saved_chunk_off[l]) * TM : nullptr;
                if (act == NVO_ACT_RELU)
                    fwd_hidden_layer<NVO_ACT_RELU>(trow, np, half, sb, sAct, saved_tile, r);
                else if (act == NVO_ACT_NONE)
                    fwd_hidden_layer<NVO_ACT_NONE>(trow, np, half, sb, sAct, saved_tile, r);
                else if (act == NVO_ACT_SIGMOID)
                    fwd_hidden_layer<NVO_ACT_SIGMOID>(trow, np, half, sb, sAct, saved_tile, r);
                else if (act == NVO_ACT_TANH)
                    fwd_hidden_layer<NVO_ACT_TANH>(trow, np, half, sb, sAct, saved_tile, r);
                else
                    fwd_hidden_layer<NVO_ACT_EXP>(trow, np, half, sb, sAct, saved_tile, r);
            } else {
                const int N = p.dims[l];
                // 16-column groups alternate between the two column halves; only the real outputs go through the activation
                for (int c16 = half * 16; c16 < N; c16 += 32) {
                    float v[16];
                    tmem_ld16(trow + c16, v);
                    if (live) {
                        const float m = row_mask ? __ldg(row_mask + row) : 1.f;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c16 + j < N) v[j] = tc_act_fwd(v[j] + sb[c16 + j], act) * m;
                        float* dst = y + row * N + c16;
                        if (N - c16 >= 16 && (N & 3) == 0) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(dst)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (c16 + j < N) dst[j] = v[j];
                        }
                    }
                }
            }
            // the next MMA (next layer or next tile) overwrites TMEM and reads the freshly written activation tile
            tc_fence_before();
            fence_async_smem();
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

// ================================================================================================================
// backward
// smem: [G0 16 KB][G1 16 KB][stage][weight image][mbar_mma, mbar_w, mbar_a[4], tmem ptr]
//   stage = for every layer l: its input tile (kpad_l/8 chunks) + ones chunk + zero chunk (the bias-gradient columns); layers whose
//           weight gradient is accumulated transposed come first (their M = 128 operand read runs 16 chunks past a_off[l])
// TMEM: [0,64) dgrad accumulator, then per layer dW_l (+ db_l): kpad_l + 16 columns (rows = output features), or, for layers with
//   <= 16 outputs, the TRANSPOSED product D[input i][output o] in 16 columns (rows = input features, row kpad_l = db).  The
//   three-layer colour head then needs 64 + 80 + 80 + 16 = 240 columns, so TWO CTAs fit one SM (256 columns each) and one CTA's
//   MMAs / bulk copies overlap the other's epilogue.
// The stage is single-buffered per layer with one mbarrier each: as soon as the commit of layer l-1's dgrad shows that layer l's
// wgrad MMAs have retired, the NEXT tile's input tile of layer l is fetched into the same buffer (bulk copy), i.e. the stage is
// refilled progressively behind the backward sweep instead of holding two whole copies.
// ================================================================================================================
__global__ void k_absmax_scale(int64_t count, const float* __restrict__ dy, float* __restrict__ scale_bits) {
    // scale_bits[0] accumulates max|dy| as an int-ordered float (non-negative)
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(__ldg(dy + i)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(scale_bits), __float_as_int(m));
}

__device__ __forceinline__ float grad_scale_from_max(float mx) {
    if (!(mx > 0.f) || !isfinite(mx)) return 1.f;
    int e;
    frexpf(mx, &e);                 // mx = f * 2^e, f in [0.5,1)
    return ldexpf(1.f, 8 - e);      // max|dy| * scale in [128, 256)
}

#define BWD_G0 0
#define BWD_G1 (8 * CHUNK_B)
#define BWD_STAGE (16 * CHUNK_B)

// bulk copy of layer l's input tile of `tile` (l == 0: the network input, else the saved activations of layer l-1)
__device__ __forceinline__ void bwd_fetch(const TcP& p, unsigned char* stage, int l, int64_t tile, const unsigned char* __restrict__ x16,
                                          const unsigned char* __restrict__ saved, uint64_t* mbar) {
    const uint32_t bytes = (uint32_t)p.kpad[l] * 256u;
    mbar_expect_tx(mbar, bytes);
    if (l == 0)
        bulk_g2s(stage + p.a_off[0], x16 + tile * p.x_tile_bytes, bytes, mbar);
    else
        bulk_g2s(stage + p.a_off[l], saved + tile * p.saved_tile_bytes + p.saved_chunk_off[l - 1] * (int64_t)CHUNK_B, bytes, mbar);
}

// dL/dz of the last layer for one (row, output o): dy * mask * act'(y)
__device__ __forceinline__ float dz_last_direct(const TcP& p, int64_t n, int64_t row, int o, const float* __restrict__ dy, const float* __restrict__ y,
                                                const float* __restrict__ row_mask) {
    const int last = p.n_layers - 1, N = p.dims[last], act = p.acts[last];
    if (row >= n || o >= N) return 0.f;
    float g = __ldg(dy + row * N + o);
    if (row_mask) g *= __ldg(row_mask + row);
    if (act != NVO_ACT_NONE) g *= tc_act_bwd(__ldg(y + row * N + o), act);
    return g;
}
// the first 16 outputs of a tile's row into registers (every field network has <= 16 real outputs); loads are issued one
// tile ahead so their latency hides behind the previous tile's layer chain
__device__ __forceinline__ void load_dz_last(const TcP& p, int64_t n, int64_t tile, int r, int64_t n_tiles, const float* __restrict__ dy,
                                             const float* __restrict__ y, const float* __restrict__ row_mask, float* pre) {
    const int last = p.n_layers - 1, N = p.dims[last], act = p.acts[last];
    const int64_t row = tile * TM + r;
    const bool live = tile < n_tiles && row < n;
    const float m = (live && row_mask) ? __ldg(row_mask + row) : 1.f;
    if (live && N == 16 && act == NVO_ACT_NONE) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(dy + row * 16) + q);
            pre[4 * q] = g.x * m, pre[4 * q + 1] = g.y * m, pre[4 * q + 2] = g.z * m, pre[4 * q + 3] = g.w * m;
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        float g = 0.f;
        if (live && j < N) {
            g = __ldg(dy + row * N + j) * m;
            if (act != NVO_ACT_NONE) g *= tc_act_bwd(__ldg(y + row * N + j), act);
        }
        pre[j] = g;
    }
}

__global__ void __launch_bounds__(NT) k_mlp_tc_bwd(const __grid_constant__ TcP p, int64_t n, const unsigned char* __restrict__ x16,
                                                   const unsigned char* __restrict__ wimg, const unsigned char* __restrict__ saved,
                                                   const float* __restrict__ y, const float* __restrict__ row_mask, const float* __restrict__ dy,
                                                   const float* __restrict__ dy_absmax, float absmax_hint, float* __restrict__ dx, float* __restrict__ dparams,
                                                   int smem_w_off, int smem_ctrl_off) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sW = smem + smem_w_off;
    unsigned char* stage = smem + BWD_STAGE;
    uint64_t* mbar_mma = reinterpret_cast<uint64_t*>(smem + smem_ctrl_off);
    uint64_t* mbar_w = mbar_mma + 1;
    uint64_t* mbar_a = mbar_mma + 2;  // [TC_MAX_LAYERS]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(mbar_mma + 2 + TC_MAX_LAYERS);
    const int tid = threadIdx.x, warp = tid >> 5, half = warp >> 2;
    const int r = ((warp & 3) << 5) | (tid & 31);  // row of the tile == TMEM lane
    const int64_t n_tiles = (n + TM - 1) / TM;
    const int last = p.n_layers - 1;
    // zero both dZ buffers and the stage once: padded chunks must hold finite values (they feed unused accumulator rows /
    // columns), then write the constant ones chunk behind every layer's input tile (the zero chunk behind it stays zero)
    for (int e = tid; e < (BWD_STAGE + p.stage_bytes) / 16; e += NT) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (tid < TM)
        for (int l = 0; l < p.n_layers; ++l)  // ones chunk: feature kpad_l == 1.0 for every row -> the dW entry of feature kpad_l accumulates sum_s dZ = db
            *reinterpret_cast<uint4*>(stage + p.a_off[l] + (p.kpad[l] >> 3) * CHUNK_B + tid * 16) = make_uint4(0x00003C00u, 0, 0, 0);
    if (warp == 0) {
        if (p.tmem_cols == 256)
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        else
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();  // generic-proxy initialisation above ordered before the async-proxy (bulk copy / MMA) accesses
    __syncthreads();
    if (tid == 0) {
        mbar_init(mbar_mma, 1);
        mbar_init(mbar_w, 1);
        for (int l = 0; l < TC_MAX_LAYERS; ++l) mbar_init(mbar_a + l, 1);
        fence_mbar_init();
        fence_async_smem();
        mbar_expect_tx(mbar_w, (uint32_t)p.img_bytes);
        bulk_g2s(sW, wimg, (uint32_t)p.img_bytes, mbar_w);
        if ((int64_t)blockIdx.x < n_tiles)
            for (int l = last; l >= 0; --l) bwd_fetch(p, stage, l, blockIdx.x, x16, saved, mbar_a + l);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const float gscale = grad_scale_from_max(absmax_hint > 0.f ? absmax_hint : __ldg(dy_absmax));
    const float inv_gscale = 1.f / gscale;
    mbar_wait(mbar_w, 0);
    uint32_t phase = 0;
    int it = 0;
    float pre[16];
    if (half == 0) load_dz_last(p, n, blockIdx.x, r, n_tiles, dy, y, row_mask, pre);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int64_t row = tile * TM + r;
        const bool live = row < n;
        const int64_t next_tile = tile + gridDim.x;
        const uint32_t a_parity = (uint32_t)(it & 1);
        // ---- dZ of the last layer: dy * mask * act'(y) * scale -> fp16 G tile (buffer parity of `last`) -------------
        // (every MMA of the previous tile has retired: its layer-0 commit sits behind the wgrad MMAs and was waited on, so both dZ
        // buffers are free.  dy, y, mask of THIS tile were loaded into registers during the previous tile; issue the next tile's now)
        {
            unsigned char* sG = smem + ((last & 1) ? BWD_G1 : BWD_G0);
            const int np = p.npad[last];
            if (half == 0) {
                float cur[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) cur[j] = pre[j];
                load_dz_last(p, n, next_tile, r, n_tiles, dy, y, row_mask, pre);
#pragma unroll
                for (int c8 = 0; c8 < 2; ++c8) {  // np >= 16: the two register-resident chunks
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = cur[c8 * 8 + j] * gscale;
                    *reinterpret_cast<uint4*>(sG + c8 * CHUNK_B + r * 16) = pack8(v);
                }
            } else {
                for (int c8 = 2; c8 < (np >> 3); ++c8) {  // wider output layers (not on the nerfacto path): direct loads
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = dz_last_direct(p, n, row, c8 * 8 + j, dy, y, row_mask) * gscale;
                    *reinterpret_cast<uint4*>(sG + c8 * CHUNK_B + r * 16) = pack8(v);
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        for (int l = last; l >= 0; --l) {
            const int kp = p.kpad[l], np = p.npad[l];
            unsigned char* sG = smem + ((l & 1) ? BWD_G1 : BWD_G0);    // dZ_l
            unsigned char* sGn = smem + ((l & 1) ? BWD_G0 : BWD_G1);   // dZ_{l-1}
            unsigned char* sA = stage + p.a_off[l];                    // input activations of layer l (+ ones / zero chunks)
            const bool need_dgrad = l > 0 || dx != nullptr;
            mbar_wait(mbar_a + l, a_parity);  // this tile's input activations of layer l have landed (read by the wgrad MMAs and by the epilogue)
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                const uint32_t g0 = smem_u32(sG), a0 = smem_u32(sA), w0 = smem_u32(sW + p.iw_off[l]);
                if (need_dgrad) {
                    // dgrad: D[s][i] = sum_o dZ[s][o] * W[o][i]; A = G tile K-major, B = W tile MN-major
                    const uint32_t idesc = umma_idesc(TM, kp, 0, 1);
                    for (int k = 0; k < (np >> 4); ++k)
                        umma_f16(tmem, umma_desc(g0 + k * 2 * CHUNK_B, CHUNK_B, 128), umma_desc(w0 + k * 256, 128, np * 16), idesc, k > 0);
                    if (l > 0) umma_commit(mbar_mma);  // the epilogue below only needs the dgrad accumulator
                }
                if (dparams) {
                    // wgrad, both operands MN-major, K = the tile's 128 samples.  Issued behind the dgrad commit: runs on the tensor pipe
                    // while the warps do the dgrad epilogue.
                    const uint32_t dcol = tmem + (uint32_t)p.dw_col[l];
                    const uint32_t acc0 = it > 0 ? 1u : 0u;
                    if (p.wg_t[l]) {
                        // narrow output layer: D[input i (M = 128, rows >= kp+16 unused)][output o (N = 16)] += sum_s Aext[s][i] * dZ[s][o]
                        const uint32_t idesc = umma_idesc(TM, 16, 1, 1);
                        for (int k = 0; k < TM / 16; ++k)
                            umma_f16(dcol, umma_desc(a0 + k * 256, 128, CHUNK_B), umma_desc(g0 + k * 256, 128, CHUNK_B), idesc, (k > 0) ? 1u : acc0);
                    } else {
                        // D[output o (M = 128 padded)][input i (N = kp+16, column kp = ones)] += sum_s dZ[s][o] * Aext[s][i]
                        const uint32_t idesc = umma_idesc(TM, kp + 16, 1, 1);
                        for (int k = 0; k < TM / 16; ++k)
                            umma_f16(dcol, umma_desc(g0 + k * 256, 128, CHUNK_B), umma_desc(a0 + k * 256, 128, CHUNK_B), idesc, (k > 0) ? 1u : acc0);
                    }
                }
                if (l == 0) umma_commit(mbar_mma);  // layer 0: the commit covers the wgrad MMAs too => the tile is fully retired
            }
            mbar_wait(mbar_mma, phase);
            phase ^= 1;
            tc_fence_after();
            // MMAs retire in order: this commit also covers layer l+1's wgrad (and, for l == 0, layer 0's), the last readers of those
            // input tiles -> refill them with the next tile's activations
            if (tid == 0 && next_tile < n_tiles) {
                if (l < last) bwd_fetch(p, stage, l + 1, next_tile, x16, saved, mbar_a + l + 1);
                if (l == 0) bwd_fetch(p, stage, 0, next_tile, x16, saved, mbar_a);
            }
            if (need_dgrad) {
                const int K = l == 0 ? p.in_dim : p.dims[l - 1];
                if (l > 0) {
                    // dZ_{l-1} = dA_{l-1} * act'(a_{l-1}); a_{l-1} is this row's entry of layer l's input tile
                    const int act = p.acts[l - 1];
                    if (act == NVO_ACT_RELU)
                        bwd_hidden_layer<NVO_ACT_RELU>(trow, kp, half, sA, sGn, r);
                    else if (act == NVO_ACT_NONE)
                        bwd_hidden_layer<NVO_ACT_NONE>(trow, kp, half, sA, sGn, r);
                    else if (act == NVO_ACT_SIGMOID)
                        bwd_hidden_layer<NVO_ACT_SIGMOID>(trow, kp, half, sA, sGn, r);
                    else if (act == NVO_ACT_TANH)
                        bwd_hidden_layer<NVO_ACT_TANH>(trow, kp, half, sA, sGn, r);
                    else
                        bwd_hidden_layer<NVO_ACT_EXP>(trow, kp, half, sA, sGn, r);
                } else {
                    int c, c1;
                    half_range(kp, half, c, c1);
                    while (c < c1) {
                        float v[32];
                        const int ncol = (c1 - c >= 32) ? 32 : 16;
                        if (ncol == 32)
                            tmem_ld32(trow + c, v);
                        else
                            tmem_ld16(trow + c, v);
                        if (live) {
                            // TMF ("tile-major float"): dx[tile][col][row] — a warp stores 128 contiguous bytes per column
                            float* dst = dx + (tile * K + c) * TM + r;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < ncol && c + j < K) dst[j * TM] = v[j] * inv_gscale;
                        }
                        c += ncol;
                    }
                }
            }
            tc_fence_before();
            fence_async_smem();
            __syncthreads();
        }
    }
    // ---- flush dW / db ------------------------------------------------------------------------------------------------------
    if (dparams && it > 0) {
        tc_fence_after();
        for (int l = 0; l < p.n_layers; ++l) {
            const int K = l == 0 ? p.in_dim : p.dims[l - 1], N = p.dims[l], kp = p.kpad[l];
            const uint32_t dcol = trow + (uint32_t)p.dw_col[l];
            if (p.wg_t[l]) {
                // lane r holds input feature r: columns = outputs (N <= 16); row kp = the ones feature = bias gradient
                if (half == 0) {
                    float v[16];
                    tmem_ld16(dcol, v);
                    if (r < K || r == kp) {
#pragma unroll
                        for (int o = 0; o < 16; ++o)
                            if (o < N) atomicAdd(r < K ? dparams + p.w_off[l] + o * K + r : dparams + p.b_off[l] + o, v[o] * inv_gscale);
                    }
                }
            } else {
                // lane r holds output feature r: columns = inputs, column kp = bias gradient; 16-column groups alternate between the halves
                for (int c16 = half * 16; c16 < kp + 16; c16 += 32) {
                    float v[16];
                    tmem_ld16(dcol + c16, v);
                    if (r < N) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int i = c16 + j;
                            if (i < K)
                                atomicAdd(dparams + p.w_off[l] + r * K + i, v[j] * inv_gscale);
                            else if (i == kp)
                                atomicAdd(dparams + p.b_off[l] + r, v[j] * inv_gscale);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        if (p.tmem_cols == 256)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// fp32 TMF [ceil(n/128)][K][128] -> row-major [n, K] (only the generic tcnn.Network wrapper needs row-major input gradients)
__global__ void __launch_bounds__(256) k_tmf_to_rows(int64_t n, int K, const float* __restrict__ src, float* __restrict__ dst) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * K) return;
    const int64_t row = t / K;
    const int c = (int)(t - row * K);
    dst[t] = __ldg(src + ((row >> 7) * K + c) * TM + (row & 127));
}

// fp32 [n, in_dim] row-major -> fp16 TMH tiles [ceil(n/128)][kpad/8][128][8], zero padded (columns >= in_dim, rows >= n)
__global__ void __launch_bounds__(256) k_cast_tmh(int64_t n, int in_dim, int kpad, const float* __restrict__ x, uint4* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one (tile, chunk, row) item per thread
    const int nch = kpad >> 3;
    const int64_t n_tiles = (n + TM - 1) / TM;
    if (t >= n_tiles * nch * TM) return;
    const int r = (int)(t & (TM - 1));
    const int64_t tc = t >> 7;
    const int c = (int)(tc % nch);
    const int64_t row = (tc / nch) * TM + r;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = c * 8 + j;
        v[j] = (row < n && col < in_dim) ? __ldg(x + row * in_dim + col) : 0.f;
    }
    out[t] = pack8(v);
}

// ================================================================================================================
static int make_tc_params(const nvo_mlp_desc* d, TcP* p) {
    NVO_CHECK(d != nullptr, "mlp_tc: null descriptor");
    NVO_CHECK(d->n_layers >= 1 && d->n_layers <= TC_MAX_LAYERS, "mlp_tc: n_layers=%d out of range [1,%d]", d->n_layers, TC_MAX_LAYERS);
    NVO_CHECK(d->in_dim >= 1 && d->in_dim <= MAXW, "mlp_tc: in_dim=%d out of range [1,%d]", d->in_dim, MAXW);
    p->n_layers = d->n_layers;
    p->in_dim = d->in_dim;
    p->k0pad = (d->in_dim + 15) & ~15;
    int off = 0, ioff = 0, chunks = 0, in = d->in_dim, aoff = 0;
    for (int l = 0; l < TC_MAX_LAYERS; ++l) {
        if (l < d->n_layers) {
            NVO_CHECK(d->dims[l] >= 1 && d->dims[l] <= MAXW, "mlp_tc: layer %d width %d out of range [1,%d]", l, d->dims[l], MAXW);
            NVO_CHECK(d->acts[l] >= NVO_ACT_NONE && d->acts[l] <= NVO_ACT_TRUNC_EXP, "mlp_tc: layer %d has unknown activation %d", l, d->acts[l]);
            p->dims[l] = d->dims[l];
            p->acts[l] = d->acts[l];
            p->npad[l] = (d->dims[l] + 15) & ~15;
            p->kpad[l] = (in + 15) & ~15;
            p->w_off[l] = off;
            off += d->dims[l] * in;
            p->b_off[l] = off;
            off += d->dims[l];
            p->iw_off[l] = ioff;
            ioff += p->npad[l] * p->kpad[l] * 2;
            p->ib_off[l] = ioff;
            ioff += p->npad[l] * 4;
            p->saved_chunk_off[l] = chunks;
            if (l < d->n_layers - 1) chunks += p->npad[l] >> 3;
            in = d->dims[l];
        } else {
            p->dims[l] = p->acts[l] = p->npad[l] = p->kpad[l] = p->w_off[l] = p->b_off[l] = p->iw_off[l] = p->ib_off[l] = p->saved_chunk_off[l] = 0;
            p->a_off[l] = 0;
        }
    }
    p->n_params = off;
    p->saved_chunks = chunks;
    p->x_tile_bytes = (int64_t)p->k0pad * 256;
    p->saved_tile_bytes = (int64_t)chunks * CHUNK_B;
    p->img_bytes = (ioff + 15) & ~15;
    // backward: weight-gradient orientation, TMEM columns and stage layout (transposed layers first, see k_mlp_tc_bwd)
    int col = MAXW;
    for (int l = 0; l < TC_MAX_LAYERS; ++l) {
        p->wg_t[l] = (l < d->n_layers && p->npad[l] == 16) ? 1 : 0;
        p->dw_col[l] = col;
        if (l < d->n_layers) col += p->wg_t[l] ? 16 : p->kpad[l] + 16;
    }
    NVO_CHECK(col <= 512, "mlp_tc: network needs %d TMEM columns (> 512)", col);
    p->tmem_cols = col <= 256 ? 256 : 512;
    aoff = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int l = 0; l < d->n_layers; ++l)
            if ((pass == 0) == (p->wg_t[l] != 0)) {
                p->a_off[l] = aoff;
                aoff += ((p->kpad[l] >> 3) + 2) * CHUNK_B;
            }
    p->stage_bytes = aoff;
    // a transposed layer's M = 128 operand read covers 16 chunks from a_off[l]: it must stay inside stage + weight image, else that
    // layer goes back to the row = output orientation (the stage order then no longer matters for it)
    bool relayout = false;
    for (int l = 0; l < d->n_layers; ++l)
        if (p->wg_t[l] && p->a_off[l] + 16 * CHUNK_B > p->stage_bytes + p->img_bytes) {
            p->wg_t[l] = 0;
            relayout = true;
        }
    if (relayout) {
        col = MAXW;
        for (int l = 0; l < d->n_layers; ++l) {
            p->dw_col[l] = col;
            col += p->wg_t[l] ? 16 : p->kpad[l] + 16;
        }
        NVO_CHECK(col <= 512, "mlp_tc: network needs %d TMEM columns (> 512)", col);
        p->tmem_cols = col <= 256 ? 256 : 512;
    }
    return 0;
}

extern "C" int64_t nvo_mlp_tc_saved_bytes(const nvo_mlp_desc* d, int64_t n) {
    TcP p;
    if (make_tc_params(d, &p)) return -1;
    const int64_t tiles = (n + TM - 1) / TM;
    return tiles * p.saved_chunks * CHUNK_B;
}

extern "C" int64_t nvo_mlp_tc_wimage_bytes(const nvo_mlp_desc* d) {
    TcP p;
    if (make_tc_params(d, &p)) return -1;
    return p.img_bytes;
}

extern "C" int nvo_mlp_tc_in_pad(const nvo_mlp_desc* d) { return d ? ((d->in_dim + 15) & ~15) : -1; }

extern "C" int nvo_cast_pad_f16(void* stream, int64_t n, int32_t in_dim, int32_t kpad, const float* x, void* out) {
    NVO_CHECK(n >= 0 && in_dim >= 1 && kpad >= in_dim && (kpad & 7) == 0, "cast_pad_f16: bad shape");
    if (n == 0) return 0;
    NVO_CHECK(x && out, "cast_pad_f16: null pointer");
    const int64_t items = ((n + TM - 1) / TM) * (kpad >> 3) * TM;
    k_cast_tmh<<<nvo_blocks(items, 256), 256, 0, (cudaStream_t)stream>>>(n, in_dim, kpad, x, (uint4*)out);
    NVO_CUDA_LAUNCH_CHECK("cast_pad_f16");
    return 0;
}

extern "C" int nvo_tmf_to_rows(void* stream, int64_t n, int32_t K, const float* src, float* dst) {
    NVO_CHECK(n >= 0 && K >= 1, "tmf_to_rows: bad shape");
    if (n == 0) return 0;
    NVO_CHECK(src && dst, "tmf_to_rows: null pointer");
    k_tmf_to_rows<<<nvo_blocks(n * K, 256), 256, 0, (cudaStream_t)stream>>>(n, K, src, dst);
    NVO_CUDA_LAUNCH_CHECK("tmf_to_rows");
    return 0;
}

extern "C" int nvo_mlp_tc_pack_weights(const nvo_mlp_desc* d, void* stream, const float* params, void* wimage) {
    TcP p;
    if (int e = make_tc_params(d, &p)) return e;
    NVO_CHECK(params && wimage, "mlp_tc_pack_weights: null pointer");
    k_tc_pack<<<8, 256, 0, (cudaStream_t)stream>>>(p, params, (unsigned char*)wimage);
    NVO_CUDA_LAUNCH_CHECK("mlp_tc_pack_weights");
    return 0;
}

extern "C" int nvo_mlp_tc_forward(const nvo_mlp_desc* d, void* stream, int64_t n, const void* x16, const void* wimage, const float* row_mask, float* y,
                                  void* saved) {
    TcP p;
    if (int e = make_tc_params(d, &p)) return e;
    NVO_CHECK(n >= 0, "mlp_tc_forward: negative batch");
    if (n == 0) return 0;
    NVO_CHECK(x16 && wimage && y, "mlp_tc_forward: null pointer");
    const int ctrl = FWD_W + p.img_bytes;
    const size_t smem = (size_t)ctrl + 48;
    cudaError_t e = cudaFuncSetAttribute(k_mlp_tc_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    NVO_CHECK(e == cudaSuccess, "mlp_tc_forward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    const int64_t tiles = (n + TM - 1) / TM;
    const int ctas_per_sm = (int)max((size_t)1, min((size_t)4, (size_t)(227 * 1024) / (smem + 1024)));
    const unsigned int grid = (unsigned int)min(tiles, (int64_t)nvo_sm_count() * ctas_per_sm);
    k_mlp_tc_fwd<<<grid, NT, smem, (cudaStream_t)stream>>>(p, n, (const unsigned char*)x16, (const unsigned char*)wimage, row_mask, y, (uint4*)saved, ctrl);
    NVO_CUDA_LAUNCH_CHECK("mlp_tc_forward");
    return 0;
}

static int mlp_tc_backward_impl(const nvo_mlp_desc* d, void* stream, int64_t n, const void* x16, int64_t x_tile_bytes, const void* wimage, const void* saved,
                                int64_t saved_tile_bytes, const float* y, const float* row_mask, const float* dy, float dy_absmax_hint, float* scratch, float* dx,
                                float* dparams) {
    TcP p;
    if (int e = make_tc_params(d, &p)) return e;
    if (x_tile_bytes > 0) p.x_tile_bytes = x_tile_bytes;
    if (saved_tile_bytes > 0) p.saved_tile_bytes = saved_tile_bytes;
    NVO_CHECK(n >= 0, "mlp_tc_backward: negative batch");
    if (n == 0) return 0;
    NVO_CHECK(x16 && wimage && dy && scratch, "mlp_tc_backward: null pointer");
    NVO_CHECK(p.n_layers == 1 || saved, "mlp_tc_backward: saved activations required for multi-layer networks");
    NVO_CHECK(p.acts[p.n_layers - 1] == NVO_ACT_NONE || y, "mlp_tc_backward: y required for an output activation");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (dy_absmax_hint == 0.f) {  // the caller does not know max|dy| (> 0: given; < 0: already in scratch): one reduction pass over dy
        e = cudaMemsetAsync(scratch, 0, sizeof(float), st);
        NVO_CHECK(e == cudaSuccess, "mlp_tc_backward: memset: %s", cudaGetErrorString(e));
        const int64_t count = n * p.dims[p.n_layers - 1];
        k_absmax_scale<<<(unsigned int)min((int64_t)nvo_sm_count() * 4, (count + 255) / 256), 256, 0, st>>>(count, dy, scratch);
        NVO_CUDA_LAUNCH_CHECK("mlp_tc_backward(absmax)");
    }
    const int w_off = BWD_STAGE + p.stage_bytes;
    const int ctrl = w_off + p.img_bytes;
    const size_t smem = (size_t)ctrl + 64;
    NVO_CHECK(smem <= 227 * 1024, "mlp_tc_backward: network needs %zu B of shared memory (> 227 KB)", smem);
    e = cudaFuncSetAttribute(k_mlp_tc_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    NVO_CHECK(e == cudaSuccess, "mlp_tc_backward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    const int64_t tiles = (n + TM - 1) / TM;
    // co-resident CTAs per SM: bounded by TMEM (512 columns) and shared memory; their MMAs / copies / epilogues overlap
    const int ctas_per_sm = (int)max((size_t)1, min((size_t)(512 / p.tmem_cols), (size_t)(227 * 1024) / (smem + 1024)));
    const unsigned int grid = (unsigned int)min(tiles, (int64_t)nvo_sm_count() * ctas_per_sm);
    k_mlp_tc_bwd<<<grid, NT, smem, st>>>(p, n, (const unsigned char*)x16, (const unsigned char*)wimage, (const unsigned char*)saved, y, row_mask, dy, scratch,
                                         dy_absmax_hint, dx, dparams, w_off, ctrl);
    NVO_CUDA_LAUNCH_CHECK("mlp_tc_backward");
    return 0;
}

extern "C" int nvo_mlp_tc_backward(const nvo_mlp_desc* d, void* stream, int64_t n, const void* x16, const void* wimage, const void* saved, const float* y,
                                   const float* row_mask, const float* dy, float dy_absmax_hint, float* scratch, float* dx, float* dparams) {
    return mlp_tc_backward_impl(d, stream, n, x16, 0, wimage, saved, 0, y, row_mask, dy, dy_absmax_hint, scratch, dx, dparams);
}

// the same with the input tiles / saved-activation tiles embedded in larger per-tile records (the fused field kernel's saved tiles):
// consecutive tiles lie x_tile_bytes / saved_tile_bytes apart
extern "C" int nvo_mlp_tc_backward_strided(const nvo_mlp_desc* d, void* stream, int64_t n, const void* x16, int64_t x_tile_bytes, const void* wimage,
                                           const void* saved, int64_t saved_tile_bytes, const float* y, const float* row_mask, const float* dy,
                                           float dy_absmax_hint, float* scratch, float* dx, float* dparams) {
    NVO_CHECK(x_tile_bytes > 0 && saved_tile_bytes > 0 && (x_tile_bytes & 15) == 0 && (saved_tile_bytes & 15) == 0, "mlp_tc_backward_strided: bad tile strides");
    return mlp_tc_backward_impl(d, stream, n, x16, x_tile_bytes, wimage, saved, saved_tile_bytes, y, row_mask, dy, dy_absmax_hint, scratch, dx, dparams);
}
