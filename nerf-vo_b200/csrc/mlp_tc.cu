// Fully fused MLP on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM) for the 64-wide field networks.
// One CTA processes 128-sample tiles; per layer ONE elected thread issues K/16 tcgen05.mma (M=128 samples, N = layer
// width, fp16 operands, fp32 accumulate), commits to an mbarrier, and the 4 warps (thread == sample row) pull the
// accumulator out of TMEM with tcgen05.ld, add bias, apply the activation, convert to fp16 and write the row straight back
// into shared memory in the UMMA canonical (no-swizzle) layout as the next layer's A operand.
//
// Shared-memory operand layout ("natural tile"): T[chunk = col/8][row][8 halfs], i.e. 16-byte vectors of 8 consecutive
// features of one row, rows adjacent at 16 B, 8-feature chunks at stride rows*16 B.  Verified on B200 (tools/umma_probe.cu):
//   * as a K-major operand (forward: A = activations, B = W[out][in]):       LBO = chunk stride, SBO = 128 B
//   * as an MN-major operand (dgrad: B = the SAME W tile; wgrad: A = dZ, B = activations, K = the 128 samples):
//                                                                             LBO = 128 B, SBO = chunk stride
// so neither transposed weight copies nor transposed activation tiles are ever materialised.
//
// Backward per layer: wgrad dW_l (+ db_l via an appended ones-column) accumulates in TMEM ACROSS all tiles of the CTA and
// is flushed once with atomics; dgrad goes TMEM -> registers -> (x relu') -> fp16 -> shared as the next dZ.  Gradients are
// scaled by a device-side power-of-two (2^8 / max|dy|) before the fp16 conversion and unscaled in fp32 on the way out.
#include <cuda_fp16.h>
#include "nvo_common.cuh"

#define TM 128           // samples per tile == threads per CTA
#define CHUNK_B 2048     // bytes of one 8-feature chunk of a 128-row tile
#define MAXW 64
#define TC_MAX_LAYERS 4
#define DW_COLS 80       // TMEM columns reserved per layer for dW (Kpad + 16 <= 80)

struct TcP {
    int n_layers, in_dim, k0pad;
    int dims[TC_MAX_LAYERS], npad[TC_MAX_LAYERS], kpad[TC_MAX_LAYERS], acts[TC_MAX_LAYERS];
    int w_off[TC_MAX_LAYERS], b_off[TC_MAX_LAYERS];      // float offsets into params
    int sw_off[TC_MAX_LAYERS], sb_off[TC_MAX_LAYERS];    // byte offsets into dynamic smem
    int saved_chunk_off[TC_MAX_LAYERS];                  // chunk offset of layer l's output inside a saved tile
    int saved_chunks;                                    // chunks per saved tile
    int n_params;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint32_t umma_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(mbar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 16 consecutive accumulator columns of this thread's row (lane) -> registers
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

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

// stage W_l (fp32 torch layout) as an fp16 natural tile with npad rows, zero padded; bias as fp32
__device__ __forceinline__ void stage_weights(const TcP& p, const float* __restrict__ params, unsigned char* smem) {
    for (int l = 0; l < p.n_layers; ++l) {
        const int K = l == 0 ? p.in_dim : p.dims[l - 1], N = p.dims[l], kp = p.kpad[l], np = p.npad[l];
        __half* sw = reinterpret_cast<__half*>(smem + p.sw_off[l]);
        float* sb = reinterpret_cast<float*>(smem + p.sb_off[l]);
        for (int e = threadIdx.x; e < np * kp; e += TM) {
            const int o = e / kp, i = e - o * kp;
            const float w = (o < N && i < K) ? __ldg(params + p.w_off[l] + o * K + i) : 0.f;
            sw[((i >> 3) * np + o) * 8 + (i & 7)] = __float2half_rn(w);
        }
        for (int o = threadIdx.x; o < np; o += TM) sb[o] = o < N ? __ldg(params + p.b_off[l] + o) : 0.f;
    }
}

// ================================================================================================================
// forward
// smem: [A tile 16 KB][W_l, b_l ...][mbar][tmem ptr]
// ================================================================================================================
__global__ void __launch_bounds__(TM) k_mlp_tc_fwd(const __grid_constant__ TcP p, int64_t n, const __half* __restrict__ x16,
                                                   const float* __restrict__ params, const float* __restrict__ row_mask, float* __restrict__ y,
                                                   __half* __restrict__ saved, int smem_ctrl_off) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + smem_ctrl_off);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + smem_ctrl_off + 8);
    const int tid = threadIdx.x, warp = tid >> 5;
    stage_weights(p, params, smem);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0;
    const int64_t n_tiles = (n + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * TM + tid;
        const bool live = row < n;
        // ---- stage the input rows (fp16, K0pad wide) as the layer-0 A operand --------------------------------
        {
            const int nch = p.k0pad >> 3;
            const uint4* src = reinterpret_cast<const uint4*>(x16 + row * p.k0pad);
            for (int c = 0; c < nch; ++c) {
                uint4 v = make_uint4(0, 0, 0, 0);
                if (live) v = __ldg(src + c);
                *reinterpret_cast<uint4*>(sA + c * CHUNK_B + tid * 16) = v;
            }
        }
        fence_async_smem();
        __syncthreads();
        for (int l = 0; l < p.n_layers; ++l) {
            const int kp = p.kpad[l], np = p.npad[l];
            if (tid == 0) {
                tc_fence_after();
                const uint32_t idesc = umma_idesc(TM, np, 0, 0);
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(smem + p.sw_off[l]);
                for (int k = 0; k < (kp >> 4); ++k)
                    umma_f16(tmem, umma_desc(a0 + k * 2 * CHUNK_B, CHUNK_B, 128), umma_desc(b0 + k * 2 * np * 16, np * 16, 128), idesc, k > 0);
                umma_commit(mbar);
            }
            mbar_wait(mbar, phase);
            phase ^= 1;
            tc_fence_after();
            const float* sb = reinterpret_cast<const float*>(smem + p.sb_off[l]);
            const bool last = l == p.n_layers - 1;
            const int act = p.acts[l];
            for (int c16 = 0; c16 < np; c16 += 16) {
                float v[16];
                tmem_ld16(trow + c16, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = tc_act_fwd(v[j] + sb[c16 + j], act);
                if (!last) {
                    const uint4 lo = pack8(v), hi = pack8(v + 8);
                    const int c8 = c16 >> 3;
                    *reinterpret_cast<uint4*>(sA + c8 * CHUNK_B + tid * 16) = lo;
                    *reinterpret_cast<uint4*>(sA + (c8 + 1) * CHUNK_B + tid * 16) = hi;
                    if (saved) {  // tile-major natural layout: fully coalesced 16-byte stores
                        uint4* dst = reinterpret_cast<uint4*>(saved) + ((tile * p.saved_chunks + p.saved_chunk_off[l] + c8) * TM + tid);
                        dst[0] = lo;
                        dst[TM] = hi;
                    }
                } else if (live) {
                    const float m = row_mask ? __ldg(row_mask + row) : 1.f;
                    const int N = p.dims[l];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c16 + j < N) y[row * N + c16 + j] = v[j] * m;
                }
            }
            // the next MMA (next layer or next tile) overwrites TMEM and reads the freshly written A tile
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
// smem: [G tile 16 KB][A tile 20 KB (8 chunks + ones chunk + zero chunk)][W_l, b_l ...][mbar][tmem ptr]
// TMEM (512 cols): [0,64) dgrad accumulator, [64 + 80*l, ...) dW_l (+ db_l in column kpad_l)
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

__global__ void __launch_bounds__(TM) k_mlp_tc_bwd(const __grid_constant__ TcP p, int64_t n, const __half* __restrict__ x16,
                                                   const float* __restrict__ params, const __half* __restrict__ saved, const float* __restrict__ y,
                                                   const float* __restrict__ row_mask, const float* __restrict__ dy, const float* __restrict__ dy_absmax,
                                                   float* __restrict__ dx, float* __restrict__ dparams, int smem_w_off, int smem_ctrl_off) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sG = smem;
    unsigned char* sA = smem + 8 * CHUNK_B;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + smem_ctrl_off);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + smem_ctrl_off + 8);
    const int tid = threadIdx.x, warp = tid >> 5;
    stage_weights(p, params, smem);
    // zero the G / A tiles once: padded chunks must hold finite values (they feed unused accumulator rows/cols)
    for (int e = tid; e < (18 * CHUNK_B) / 16; e += TM) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0, 0, 0, 0);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const float gscale = grad_scale_from_max(__ldg(dy_absmax));
    const float inv_gscale = 1.f / gscale;
    uint32_t phase = 0;
    const int last = p.n_layers - 1;
    const int64_t n_tiles = (n + TM - 1) / TM;
    int tiles_done = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tiles_done) {
        const int64_t row = tile * TM + tid;
        const bool live = row < n;
        // ---- dZ of the last layer: dy * mask * act'(y) * scale -> fp16 G tile ----------------------------------
        {
            const int N = p.dims[last], np = p.npad[last], act = p.acts[last];
            const float m = (live && row_mask) ? __ldg(row_mask + row) : 1.f;
            for (int c8 = 0; c8 < (np >> 3); ++c8) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int o = c8 * 8 + j;
                    float g = 0.f;
                    if (live && o < N) {
                        g = __ldg(dy + row * N + o) * m;
                        if (act != NVO_ACT_NONE) {
                            float a = __ldg(y + row * N + o);
                            g *= tc_act_bwd(a, act);
                        }
                    }
                    v[j] = g * gscale;
                }
                *reinterpret_cast<uint4*>(sG + c8 * CHUNK_B + tid * 16) = pack8(v);
            }
        }
        for (int l = last; l >= 0; --l) {
            const int kp = p.kpad[l], np = p.npad[l];
            // ---- stage the layer's input activations A_{l-1} (+ ones / zero chunks for the bias gradient) -------
            {
                const int nch = kp >> 3;
                if (l == 0) {
                    const uint4* src = reinterpret_cast<const uint4*>(x16 + row * p.k0pad);
                    for (int c = 0; c < nch; ++c) {
                        uint4 v = make_uint4(0, 0, 0, 0);
                        if (live) v = __ldg(src + c);
                        *reinterpret_cast<uint4*>(sA + c * CHUNK_B + tid * 16) = v;
                    }
                } else {
                    const uint4* src = reinterpret_cast<const uint4*>(saved) + ((tile * p.saved_chunks + p.saved_chunk_off[l - 1]) * TM + tid);
                    for (int c = 0; c < nch; ++c) {
                        uint4 v = __ldg(src + c * TM);
                        if (!live) v = make_uint4(0, 0, 0, 0);
                        *reinterpret_cast<uint4*>(sA + c * CHUNK_B + tid * 16) = v;
                    }
                }
                // ones chunk: feature kp == 1.0 for every row  ->  dW column kp accumulates sum_s dZ = db
                *reinterpret_cast<uint4*>(sA + nch * CHUNK_B + tid * 16) = make_uint4(0x00003C00u, 0, 0, 0);
                *reinterpret_cast<uint4*>(sA + (nch + 1) * CHUNK_B + tid * 16) = make_uint4(0, 0, 0, 0);
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            const bool need_dgrad = l > 0 || dx != nullptr;
            if (tid == 0) {
                tc_fence_after();
                const uint32_t g0 = smem_u32(sG), a0 = smem_u32(sA), w0 = smem_u32(smem + p.sw_off[l]);
                if (dparams) {
                    // wgrad: D[feature o (M=128 padded)][input i (N = kp+16)] += sum_s dZ[s][o] * A[s][i]; both operands MN-major, K = samples
                    const uint32_t idesc = umma_idesc(TM, kp + 16, 1, 1);
                    const uint32_t dcol = tmem + 64 + DW_COLS * l;
                    for (int k = 0; k < TM / 16; ++k)
                        umma_f16(dcol, umma_desc(g0 + k * 256, 128, CHUNK_B), umma_desc(a0 + k * 256, 128, CHUNK_B), idesc, (tiles_done > 0 || k > 0) ? 1u : 0u);
                }
                if (need_dgrad) {
                    // dgrad: D[s][i] = sum_o dZ[s][o] * W[o][i]; A = G tile K-major, B = W tile MN-major
                    const uint32_t idesc = umma_idesc(TM, kp, 0, 1);
                    for (int k = 0; k < (np >> 4); ++k)
                        umma_f16(tmem, umma_desc(g0 + k * 2 * CHUNK_B, CHUNK_B, 128), umma_desc(w0 + k * 256, 128, np * 16), idesc, k > 0);
                }
                umma_commit(mbar);
            }
            mbar_wait(mbar, phase);
            phase ^= 1;
            tc_fence_after();
            if (need_dgrad) {
                const int K = l == 0 ? p.in_dim : p.dims[l - 1];
                for (int c16 = 0; c16 < kp; c16 += 16) {
                    float v[16];
                    tmem_ld16(trow + c16, v);
                    if (l > 0) {
                        // dZ_{l-1} = dA_{l-1} * act'(a_{l-1}); a_{l-1} is this row's entry of the A tile
                        const int act = p.acts[l - 1];
                        const int c8 = c16 >> 3;
                        const uint4 a_lo = *reinterpret_cast<const uint4*>(sA + c8 * CHUNK_B + tid * 16);
                        const uint4 a_hi = *reinterpret_cast<const uint4*>(sA + (c8 + 1) * CHUNK_B + tid * 16);
                        const __half2* h_lo = reinterpret_cast<const __half2*>(&a_lo);
                        const __half2* h_hi = reinterpret_cast<const __half2*>(&a_hi);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f0 = __half22float2(h_lo[j]), f1 = __half22float2(h_hi[j]);
                            v[2 * j] *= tc_act_bwd(f0.x, act);
                            v[2 * j + 1] *= tc_act_bwd(f0.y, act);
                            v[8 + 2 * j] *= tc_act_bwd(f1.x, act);
                            v[8 + 2 * j + 1] *= tc_act_bwd(f1.y, act);
                        }
                        *reinterpret_cast<uint4*>(sG + c8 * CHUNK_B + tid * 16) = pack8(v);
                        *reinterpret_cast<uint4*>(sG + (c8 + 1) * CHUNK_B + tid * 16) = pack8(v + 8);
                    } else if (live) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c16 + j < K) dx[row * K + c16 + j] = v[j] * inv_gscale;
                    }
                }
            }
            tc_fence_before();
            fence_async_smem();
            __syncthreads();
        }
    }
    // ---- flush dW / db: thread o holds row o of every layer's accumulator ------------------------------------------
    if (dparams && tiles_done > 0) {
        tc_fence_after();
        for (int l = 0; l < p.n_layers; ++l) {
            const int K = l == 0 ? p.in_dim : p.dims[l - 1], N = p.dims[l], kp = p.kpad[l];
            for (int c16 = 0; c16 < kp + 16; c16 += 16) {
                float v[16];
                tmem_ld16(trow + 64 + DW_COLS * l + c16, v);
                if (tid < N) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int i = c16 + j;
                        if (i < K)
                            atomicAdd(dparams + p.w_off[l] + tid * K + i, v[j] * inv_gscale);
                        else if (i == kp)
                            atomicAdd(dparams + p.b_off[l] + tid, v[j] * inv_gscale);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// fp32 [n, in_dim] -> fp16 [n, kpad] zero padded (generic entry for tcnn.Network inputs)
__global__ void __launch_bounds__(256) k_cast_pad(int64_t n, int in_dim, int kpad, const float* __restrict__ x, __half* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * kpad) return;
    const int64_t r = t / kpad;
    const int c = (int)(t - r * kpad);
    out[t] = __float2half_rn(c < in_dim ? __ldg(x + r * in_dim + c) : 0.f);
}

// ================================================================================================================
static int make_tc_params(const nvo_mlp_desc* d, TcP* p, int base_off, int* w_region_end) {
    NVO_CHECK(d != nullptr, "mlp_tc: null descriptor");
    NVO_CHECK(d->n_layers >= 1 && d->n_layers <= TC_MAX_LAYERS, "mlp_tc: n_layers=%d out of range [1,%d]", d->n_layers, TC_MAX_LAYERS);
    NVO_CHECK(d->in_dim >= 1 && d->in_dim <= MAXW, "mlp_tc: in_dim=%d out of range [1,%d]", d->in_dim, MAXW);
    p->n_layers = d->n_layers;
    p->in_dim = d->in_dim;
    p->k0pad = (d->in_dim + 15) & ~15;
    int off = 0, soff = base_off, chunks = 0, in = d->in_dim;
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
            p->sw_off[l] = soff;
            soff += p->npad[l] * p->kpad[l] * 2;
            p->sb_off[l] = soff;
            soff += p->npad[l] * 4;
            p->saved_chunk_off[l] = chunks;
            if (l < d->n_layers - 1) chunks += p->npad[l] >> 3;
            in = d->dims[l];
        } else {
            p->dims[l] = p->acts[l] = p->npad[l] = p->kpad[l] = p->w_off[l] = p->b_off[l] = p->sw_off[l] = p->sb_off[l] = p->saved_chunk_off[l] = 0;
        }
    }
    p->n_params = off;
    p->saved_chunks = chunks;
    *w_region_end = (soff + 15) & ~15;
    return 0;
}

extern "C" int64_t nvo_mlp_tc_saved_bytes(const nvo_mlp_desc* d, int64_t n) {
    TcP p;
    int e;
    if (make_tc_params(d, &p, 0, &e)) return -1;
    const int64_t tiles = (n + TM - 1) / TM;
    return tiles * p.saved_chunks * CHUNK_B;
}

extern "C" int nvo_mlp_tc_in_pad(const nvo_mlp_desc* d) { return d ? ((d->in_dim + 15) & ~15) : -1; }

extern "C" int nvo_cast_pad_f16(void* stream, int64_t n, int32_t in_dim, int32_t kpad, const float* x, void* out) {
    NVO_CHECK(n >= 0 && in_dim >= 1 && kpad >= in_dim, "cast_pad_f16: bad shape");
    if (n == 0) return 0;
    NVO_CHECK(x && out, "cast_pad_f16: null pointer");
    k_cast_pad<<<nvo_blocks(n * kpad, 256), 256, 0, (cudaStream_t)stream>>>(n, in_dim, kpad, x, (__half*)out);
    NVO_CUDA_LAUNCH_CHECK("cast_pad_f16");
    return 0;
}

extern "C" int nvo_mlp_tc_forward(const nvo_mlp_desc* d, void* stream, int64_t n, const void* x16, const float* params, const float* row_mask, float* y,
                                  void* saved) {
    TcP p;
    int wend;
    if (int e = make_tc_params(d, &p, 8 * CHUNK_B, &wend)) return e;
    NVO_CHECK(n >= 0, "mlp_tc_forward: negative batch");
    if (n == 0) return 0;
    NVO_CHECK(x16 && params && y, "mlp_tc_forward: null pointer");
    const int ctrl = wend;
    const size_t smem = (size_t)ctrl + 16;
    cudaError_t e = cudaFuncSetAttribute(k_mlp_tc_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    NVO_CHECK(e == cudaSuccess, "mlp_tc_forward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    const int64_t tiles = (n + TM - 1) / TM;
    const unsigned int grid = (unsigned int)min(tiles, (int64_t)nvo_sm_count() * 4);
    k_mlp_tc_fwd<<<grid, TM, smem, (cudaStream_t)stream>>>(p, n, (const __half*)x16, params, row_mask, y, (__half*)saved, ctrl);
    NVO_CUDA_LAUNCH_CHECK("mlp_tc_forward");
    return 0;
}

extern "C" int nvo_mlp_tc_backward(const nvo_mlp_desc* d, void* stream, int64_t n, const void* x16, const float* params, const void* saved, const float* y,
                                   const float* row_mask, const float* dy, float* scratch, float* dx, float* dparams) {
    TcP p;
    int wend;
    if (int e = make_tc_params(d, &p, 18 * CHUNK_B, &wend)) return e;
    NVO_CHECK(n >= 0, "mlp_tc_backward: negative batch");
    if (n == 0) return 0;
    NVO_CHECK(x16 && params && dy && scratch, "mlp_tc_backward: null pointer");
    NVO_CHECK(p.n_layers == 1 || saved, "mlp_tc_backward: saved activations required for multi-layer networks");
    NVO_CHECK(p.acts[p.n_layers - 1] == NVO_ACT_NONE || y, "mlp_tc_backward: y required for an output activation");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(float), st);
    NVO_CHECK(e == cudaSuccess, "mlp_tc_backward: memset: %s", cudaGetErrorString(e));
    const int64_t count = n * p.dims[p.n_layers - 1];
    k_absmax_scale<<<(unsigned int)min((int64_t)nvo_sm_count() * 4, (count + 255) / 256), 256, 0, st>>>(count, dy, scratch);
    NVO_CUDA_LAUNCH_CHECK("mlp_tc_backward(absmax)");
    const int ctrl = wend;
    const size_t smem = (size_t)ctrl + 16;
    e = cudaFuncSetAttribute(k_mlp_tc_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    NVO_CHECK(e == cudaSuccess, "mlp_tc_backward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    const int64_t tiles = (n + TM - 1) / TM;
    const unsigned int grid = (unsigned int)min(tiles, (int64_t)nvo_sm_count());
    k_mlp_tc_bwd<<<grid, TM, smem, st>>>(p, n, (const __half*)x16, params, (const __half*)saved, y, row_mask, dy, scratch, dx, dparams, 18 * CHUNK_B, ctrl);
    NVO_CUDA_LAUNCH_CHECK("mlp_tc_backward");
    return 0;
}
