// Mapping-step losses and their gradients w.r.t. the per-sample weights (NS/model_components/losses.py).
// One warp per ray, per-ray vectors in shared memory; batch means are accumulated with one atomicAdd per CTA.
#include "nvo_common.cuh"

#define RAYS_PER_BLOCK 4
#define LOSS_EPS 1.0e-7f  // losses.py:37
#define MAX_S 1024

__device__ __forceinline__ void block_accumulate(float v_lane0, float* loss, float* red) {
    // v_lane0: per-warp value valid in lane 0.  red: RAYS_PER_BLOCK floats of smem.
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) red[wid] = v_lane0;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < RAYS_PER_BLOCK; ++i) s += red[i];
        atomicAdd(loss, s);
    }
}

// ---- distortion (losses.py:134-153) ------------------------------------------------------------------------
// smem per warp: w[S], m[S]
template <bool BWD>
__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK) k_distortion(int64_t B, int S, const float* __restrict__ weights, const float* __restrict__ sdist,
                                                                    const float* __restrict__ dscale, float scale, float* __restrict__ loss, float* __restrict__ dweights) {
    if (BWD && dscale) scale *= __ldg(dscale);
    extern __shared__ float smf[];
    __shared__ float red[RAYS_PER_BLOCK];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * RAYS_PER_BLOCK + wid;
    const bool live = r < B;
    float* w = smf + wid * 2 * S;
    float* m = w + S;
    float ray_loss = 0.f;
    if (live) {
        const float* sd = sdist + r * (S + 1);
        for (int i = lane; i < S; i += 32) {
            w[i] = __ldg(weights + r * S + i);
            m[i] = (__ldg(sd + i + 1) + __ldg(sd + i)) / 2.f;
        }
        __syncwarp();
        for (int i = lane; i < S; i += 32) {
            const float mi = m[i], wi = w[i];
            float inner = 0.f;
            for (int j = 0; j < S; ++j) inner += w[j] * fabsf(mi - m[j]);
            const float delta = __ldg(sd + i + 1) - __ldg(sd + i);
            if (BWD)
                dweights[r * S + i] += scale * (2.f * inner + 2.f * wi * delta / 3.f);
            else
                ray_loss += wi * inner + wi * wi * delta / 3.f;
        }
    }
    if (!BWD) {
        ray_loss = nvo_warp_sum(ray_loss);
        block_accumulate(live ? ray_loss * scale : 0.f, loss, red);
    }
}

// ---- interlevel (losses.py:52-130), one proposal level ----------------------------------------------------
// smem per warp: cy1[Sp+1], cp[Sp+1], E[Sp+1]
template <bool BWD>
__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK)
    k_interlevel(int64_t B, int S, int Sp, const float* __restrict__ w, const float* __restrict__ c, const float* __restrict__ wp, const float* __restrict__ cp,
                 const float* __restrict__ dscale, float scale, float* __restrict__ loss, int32_t* __restrict__ idx_lo_out, int32_t* __restrict__ idx_hi_out,
                 float* __restrict__ dwp) {
    if (BWD && dscale) scale *= __ldg(dscale);
    extern __shared__ float smf[];
    __shared__ float red[RAYS_PER_BLOCK];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * RAYS_PER_BLOCK + wid;
    const bool live = r < B;
    const int np = Sp + 1;
    float* cy1 = smf + wid * 3 * np;
    float* cps = cy1 + np;
    float* E = cps + np;
    float ray_loss = 0.f;
    if (live) {
        // cy1 = [0, cumsum(wp)]
        double carry = 0.0;
        for (int c0 = 0; c0 < Sp; c0 += 32) {
            const int i = c0 + lane;
            const float v = i < Sp ? __ldg(wp + r * Sp + i) : 0.f;
            const double incl = nvo_warp_scan_incl((double)v, lane);
            if (i < Sp) cy1[i + 1] = (float)(carry + incl);
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) cy1[0] = 0.f;
        for (int i = lane; i < np; i += 32) {
            cps[i] = __ldg(cp + r * np + i);
            if (BWD) E[i] = 0.f;
        }
        __syncwarp();
        for (int i = lane; i < S; i += 32) {
            const float t0s = __ldg(c + r * (S + 1) + i), t0e = __ldg(c + r * (S + 1) + i + 1);
            // idx_lo = searchsorted(cp[:-1], t0s, right) - 1 ; idx_hi = searchsorted(cp[1:], t0e, right); both clamped to [0, Sp-1]
            int lo = 0, hi = Sp;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cps[mid] <= t0s) lo = mid + 1; else hi = mid;
            }
            const int idx_lo = min(max(lo - 1, 0), Sp - 1);
            lo = 0, hi = Sp;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cps[mid + 1] <= t0e) lo = mid + 1; else hi = mid;
            }
            const int idx_hi = min(max(lo, 0), Sp - 1);
            if (!BWD && idx_lo_out) idx_lo_out[r * S + i] = idx_lo;
            if (!BWD && idx_hi_out) idx_hi_out[r * S + i] = idx_hi;
            const float w_outer = cy1[idx_hi + 1] - cy1[idx_lo];
            const float wi = __ldg(w + r * S + i);
            const float d = fmaxf(wi - w_outer, 0.f);
            if (BWD) {
                const float g = -2.f * d / (wi + LOSS_EPS) * scale;  // dL/d w_outer
                if (g != 0.f) {
                    atomicAdd(E + idx_hi + 1, g);
                    atomicAdd(E + idx_lo, -g);
                }
            } else {
                ray_loss += d * d / (wi + LOSS_EPS);
            }
        }
        if (BWD) {
            __syncwarp();
            // dL/dwp_j = sum_{k>j} E[k]  (cy1[k] contains wp_j iff j < k)
            float tot = 0.f;
            for (int i = lane; i < np; i += 32) tot += E[i];
            tot = nvo_warp_sum(tot);
            float carry_f = 0.f;
            for (int c0 = 0; c0 < Sp; c0 += 32) {
                const int j = c0 + lane;
                const float v = j < Sp ? E[j] : 0.f;
                const float incl = nvo_warp_scan_incl(v, lane);  // sum_{k<=j}
                if (j < Sp) dwp[r * Sp + j] += tot - (carry_f + incl);
                carry_f += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
    }
    if (!BWD) {
        ray_loss = nvo_warp_sum(ray_loss);
        block_accumulate(live ? ray_loss * scale : 0.f, loss, red);
    }
}

// ---- DS-NeRF depth loss (losses.py:224-246) ----------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK)
    k_depth_loss(int64_t B, int S, const float* __restrict__ weights, const float* __restrict__ starts, const float* __restrict__ ends, int64_t stride,
                 const float* __restrict__ depth_gt,
                 const float* __restrict__ dnorm, float sigma, const float* __restrict__ dscale, float scale, float* __restrict__ loss,
                 float* __restrict__ dweights) {
    if (BWD && dscale) scale *= __ldg(dscale);
    __shared__ float red[RAYS_PER_BLOCK];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * RAYS_PER_BLOCK + wid;
    const bool live = r < B;
    float ray_loss = 0.f;
    if (live) {
        const float D = __ldg(depth_gt + r) * __ldg(dnorm + r);  // losses.py:313-314 (is_euclidean=False)
        const bool mask = D > 0.f;
        const float two_sigma = 2.f * sigma;  // losses.py:243: (2 * sigma), not 2*sigma^2
        const float* st = starts + r * stride;
        const float* en = ends + r * stride;
        if (mask || !BWD) {
            for (int i = lane; i < S; i += 32) {
                const float s0 = __ldg(st + i), s1 = __ldg(en + i);
                const float t = (s0 + s1) / 2.f, len = s1 - s0;
                const float diff = t - D;
                const float e = expf(-(diff * diff) / two_sigma);
                const float wi = __ldg(weights + r * S + i);
                if (BWD) {
                    if (e != 0.f) dweights[r * S + i] += scale * (-1.f / (wi + LOSS_EPS)) * e * len;
                } else {
                    ray_loss += -logf(wi + LOSS_EPS) * e * len;
                }
            }
        }
        if (!mask) ray_loss = 0.f;
    }
    if (!BWD) {
        ray_loss = nvo_warp_sum(ray_loss);
        block_accumulate(live ? ray_loss * scale : 0.f, loss, red);
    }
}

// ---- MSE and MonoSDF normal loss on [B,3] maps ----------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mse(int64_t n, const float* __restrict__ pred, const float* __restrict__ target, float inv_n, float scale,
                                             float* __restrict__ loss, float* __restrict__ d_pred) {
    __shared__ float red[8];
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float v = 0.f;
    if (t < n) {
        const float d = __ldg(pred + t) - __ldg(target + t);
        v = d * d * inv_n;
        if (d_pred) d_pred[t] = scale * 2.f * d * inv_n;
    }
    v = nvo_warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0 && loss) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        atomicAdd(loss, s);
    }
}

__global__ void __launch_bounds__(256) k_normal_loss(int64_t B, const float* __restrict__ pred, const float* __restrict__ gt, float inv_B, float scale,
                                                     float* __restrict__ loss, float* __restrict__ d_pred) {
    __shared__ float red[8];
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float v = 0.f;
    if (r < B) {
        float p[3], g[3];
        float np = 0.f, ng = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            p[a] = __ldg(pred + 3 * r + a);
            g[a] = __ldg(gt + 3 * r + a);
            np += p[a] * p[a];
            ng += g[a] * g[a];
        }
        np = fmaxf(sqrtf(np), 1e-12f);  // F.normalize eps
        ng = fmaxf(sqrtf(ng), 1e-12f);
        float dot = 0.f, l1 = 0.f, dp[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            p[a] /= np;
            g[a] /= ng;
            dot += p[a] * g[a];
            const float d = p[a] - g[a];
            l1 += fabsf(d);
            dp[a] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) - g[a];  // d(l1 + 1 - p.g)/dp
        }
        v = (l1 + 1.f - dot) * inv_B;
        if (d_pred) {
            const float pd = p[0] * dp[0] + p[1] * dp[1] + p[2] * dp[2];
#pragma unroll
            for (int a = 0; a < 3; ++a) d_pred[3 * r + a] = scale * inv_B * (dp[a] - p[a] * pd) / np;
        }
    }
    v = nvo_warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0 && loss) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        atomicAdd(loss, s);
    }
}

// ---- C ABI ---------------------------------------------------------------------------------------------------
static inline unsigned int ray_blocks(int64_t B) { return (unsigned int)((B + RAYS_PER_BLOCK - 1) / RAYS_PER_BLOCK); }

extern "C" int nvo_distortion_loss_forward(void* stream, int64_t B, int32_t S, const float* weights, const float* sdist, float* loss) {
    NVO_CHECK(B >= 0 && S >= 1 && S <= MAX_S, "distortion_loss_forward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(weights && sdist && loss, "distortion_loss_forward: null pointer");
    k_distortion<false><<<ray_blocks(B), 32 * RAYS_PER_BLOCK, sizeof(float) * 2 * S * RAYS_PER_BLOCK, (cudaStream_t)stream>>>(B, S, weights, sdist, nullptr, 1.f / (float)B,
                                                                                                                       loss, nullptr);
    NVO_CUDA_LAUNCH_CHECK("distortion_loss_forward");
    return 0;
}
extern "C" int nvo_distortion_loss_backward(void* stream, int64_t B, int32_t S, const float* weights, const float* sdist, const float* dscale, float scale,
                                            float* dweights) {
    NVO_CHECK(B >= 0 && S >= 1 && S <= MAX_S, "distortion_loss_backward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(weights && sdist && dweights, "distortion_loss_backward: null pointer");
    k_distortion<true><<<ray_blocks(B), 32 * RAYS_PER_BLOCK, sizeof(float) * 2 * S * RAYS_PER_BLOCK, (cudaStream_t)stream>>>(B, S, weights, sdist, dscale, scale / (float)B,
                                                                                                                      nullptr, dweights);
    NVO_CUDA_LAUNCH_CHECK("distortion_loss_backward");
    return 0;
}

extern "C" int nvo_interlevel_loss_forward(void* stream, int64_t B, int32_t S, int32_t Sp, const float* w, const float* c, const float* wp, const float* cp,
                                           float* loss, int32_t* idx_lo, int32_t* idx_hi) {
    NVO_CHECK(B >= 0 && S >= 1 && Sp >= 1 && Sp <= MAX_S, "interlevel_loss_forward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(w && c && wp && cp && loss, "interlevel_loss_forward: null pointer");
    k_interlevel<false><<<ray_blocks(B), 32 * RAYS_PER_BLOCK, sizeof(float) * 3 * (Sp + 1) * RAYS_PER_BLOCK, (cudaStream_t)stream>>>(
        B, S, Sp, w, c, wp, cp, nullptr, 1.f / ((float)B * (float)S), loss, idx_lo, idx_hi, nullptr);
    NVO_CUDA_LAUNCH_CHECK("interlevel_loss_forward");
    return 0;
}
extern "C" int nvo_interlevel_loss_backward(void* stream, int64_t B, int32_t S, int32_t Sp, const float* w, const float* c, const float* wp, const float* cp,
                                            const float* dscale, float scale, float* dwp) {
    NVO_CHECK(B >= 0 && S >= 1 && Sp >= 1 && Sp <= MAX_S, "interlevel_loss_backward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(w && c && wp && cp && dwp, "interlevel_loss_backward: null pointer");
    k_interlevel<true><<<ray_blocks(B), 32 * RAYS_PER_BLOCK, sizeof(float) * 3 * (Sp + 1) * RAYS_PER_BLOCK, (cudaStream_t)stream>>>(
        B, S, Sp, w, c, wp, cp, dscale, scale / ((float)B * (float)S), nullptr, nullptr, nullptr, dwp);
    NVO_CUDA_LAUNCH_CHECK("interlevel_loss_backward");
    return 0;
}

extern "C" int nvo_depth_loss_forward(void* stream, int64_t B, int32_t S, const float* weights, const float* starts, const float* ends, int64_t stride,
                                      const float* depth_gt,
                                      const float* directions_norm, float sigma, float* loss) {
    NVO_CHECK(B >= 0 && S >= 1, "depth_loss_forward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(weights && starts && ends && depth_gt && directions_norm && loss, "depth_loss_forward: null pointer");
    k_depth_loss<false><<<ray_blocks(B), 32 * RAYS_PER_BLOCK, 0, (cudaStream_t)stream>>>(B, S, weights, starts, ends, stride, depth_gt, directions_norm, sigma, nullptr, 1.f / (float)B, loss,
                                                                                        nullptr);
    NVO_CUDA_LAUNCH_CHECK("depth_loss_forward");
    return 0;
}
extern "C" int nvo_depth_loss_backward(void* stream, int64_t B, int32_t S, const float* weights, const float* starts, const float* ends, int64_t stride,
                                       const float* depth_gt,
                                       const float* directions_norm, float sigma, const float* dscale, float scale, float* dweights) {
    NVO_CHECK(B >= 0 && S >= 1, "depth_loss_backward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(weights && starts && ends && depth_gt && directions_norm && dweights, "depth_loss_backward: null pointer");
    k_depth_loss<true><<<ray_blocks(B), 32 * RAYS_PER_BLOCK, 0, (cudaStream_t)stream>>>(B, S, weights, starts, ends, stride, depth_gt, directions_norm, sigma, dscale, scale / (float)B,
                                                                                       nullptr, dweights);
    NVO_CUDA_LAUNCH_CHECK("depth_loss_backward");
    return 0;
}

extern "C" int nvo_mse_loss(void* stream, int64_t n, const float* pred, const float* target, float scale, float* loss, float* d_pred) {
    NVO_CHECK(n >= 0, "mse_loss: bad shape");
    if (n == 0) return 0;
    NVO_CHECK(pred && target, "mse_loss: null pointer");
    k_mse<<<nvo_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(n, pred, target, 1.f / (float)n, scale, loss, d_pred);
    NVO_CUDA_LAUNCH_CHECK("mse_loss");
    return 0;
}
extern "C" int nvo_normal_loss(void* stream, int64_t B, const float* pred, const float* gt, float scale, float* loss, float* d_pred) {
    NVO_CHECK(B >= 0, "normal_loss: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(pred && gt, "normal_loss: null pointer");
    k_normal_loss<<<nvo_blocks(B, 256), 256, 0, (cudaStream_t)stream>>>(B, pred, gt, 1.f / (float)B, scale, loss, d_pred);
    NVO_CUDA_LAUNCH_CHECK("normal_loss");
    return 0;
}

// ================================================================================================================================
// The whole loss wave of one mapping step in ONE launch (the trainer's path: the step differentiates the total with grad_output == 1):
// every term above plus its gradient w.r.t. the three weight sets, the rendered colour and the rendered normals, per ray by one warp,
// and the weighted total by the last CTA to finish.  Same per-ray arithmetic, in the same order, as the kernels above (the two levels'
// gradient buffers receive interlevel + depth in that order); what disappears is 14 launches spread over four streams, three zero
// fills and the cuBLAS dot of the five terms, i.e. ~60 us of launch latency between the renderer and the backward pass.
// ================================================================================================================================
struct StepLossP {
    int64_t B;
    int S[3];                  // samples per level (proposal 0, proposal 1, final)
    const float* w[3];         // [B,S_l]
    const float* sdist[3];     // [B,S_l+1] normalised bin edges
    const float* starts[3];    // Euclidean interval starts / ends (row stride `stride[l]`), depth loss
    const float* ends[3];
    int64_t stride[3];
    const float *rgb, *rgb_gt;           // [B,3]
    const float *normals, *normal_gt;    // [B,3] or null
    const float *depth_gt, *dnorm;       // [B] or null
    float sigma;
    float mults[5];            // rgb, interlevel, distortion, depth (per level), normal
    float* terms;              // [5] unweighted batch means, accumulated (zeroed by the caller)
    float* total;              // [1] written by the last CTA
    unsigned int* ticket;      // [1] zero before the launch; reset by the last CTA
    float* dw[3];              // [B,S_l] written
    float* d_rgb;              // [B,3] written
    float* d_normals;          // [B,3] written, or null
};

// One proposal level's interlevel + depth terms for one ray, by one warp — the lane-blocked variant: lane j owns the C CONSECUTIVE samples
// [C j, C j + C) (C = ceil(S_l / 32)), so the weight cumsum and the gradient's suffix sum are C serial adds + ONE warp scan of the lane totals
// instead of one warp scan per 32-sample chunk (8 dependent shuffle chains at 256 samples), the depth term's expf / logf run as C independent
// chains per lane, and the two bin searches of a query advance in the same loop.  Same terms as the chunked path below; the cumsum is
// still accumulated in double and rounded to fp32 per entry.  Shared rows indexed by sample are padded (PI) so that the lane-blocked
// accesses of an even C hit 32 different banks.
template <int C>
__device__ __forceinline__ int sl_pi(int i) { return (C % 2 == 0) ? i + i / C : i; }
template <int C>
__device__ __forceinline__ void sl_level_fast(const StepLossP& p, const int l, const int64_t r, const int lane, const float* __restrict__ wf, const float* __restrict__ cf,
                                              float* __restrict__ cy1, float* __restrict__ cps, float* __restrict__ E, const float D, const bool use_depth,
                                              const bool dmask, const float two_sigma, const float g_depth, const float B_f, float& t_inter, float& t_depth) {
    const int Sp = p.S[l], np = Sp + 1, S = p.S[2];
    const float* wp = p.w[l] + r * Sp;
    const int j0 = C * lane;
    float v[C];
    if (C % 4 == 0 && j0 + C <= Sp && (reinterpret_cast<uintptr_t>(wp) & 15) == 0) {
#pragma unroll
        for (int k = 0; k < C; k += 4) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(wp + j0 + k));
            v[k] = q.x, v[k + 1] = q.y, v[k + 2] = q.z, v[k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < C; ++k) v[k] = j0 + k < Sp ? __ldg(wp + j0 + k) : 0.f;
    }
    // interval geometry of the lane's samples (depth term): requested now, consumed at the end
    float s0[C], s1[C];
    if (use_depth) {
        const float* st = p.starts[l] + r * p.stride[l];
        const float* en = p.ends[l] + r * p.stride[l];
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const bool in = j0 + k < Sp;
            s0[k] = in ? __ldg(st + j0 + k) : 0.f;
            s1[k] = in ? __ldg(en + j0 + k) : 0.f;
        }
    }
    double loc[C], s = 0.0;
#pragma unroll
    for (int k = 0; k < C; ++k) s += (double)v[k], loc[k] = s;
    const double incl = nvo_warp_scan_incl(s, lane);
    double excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 0.0;
#pragma unroll
    for (int k = 0; k < C; ++k)
        if (j0 + k < Sp) cy1[sl_pi<C>(j0 + k + 1)] = (float)(excl + loc[k]);
    if (lane == 0) cy1[0] = 0.f;
    for (int i = lane; i < np; i += 32) cps[i] = __ldg(p.sdist[l] + r * np + i);
    const int np_pad = sl_pi<C>(np) + 1;
    for (int i = lane; i < np_pad; i += 32) E[i] = 0.f;
    __syncwarp();
    const float g_scale = p.mults[1] / (B_f * (float)S);
    float ray_loss = 0.f;
    // two field samples per lane and trip (S = 48: one trip), four bin searches advancing together
    for (int i0 = lane; i0 < S; i0 += 64) {
        float ts[2], te[2];
        int lo[2], hi[2], lo2[2], hi2[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = i0 + 32 * q;
            const bool act = i < S;
            ts[q] = act ? cf[i] : 0.f, te[q] = act ? cf[i + 1] : 0.f;
            lo[q] = lo2[q] = 0, hi[q] = hi2[q] = act ? Sp : 0;
        }
        while (lo[0] < hi[0] || lo2[0] < hi2[0] || lo[1] < hi[1] || lo2[1] < hi2[1]) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (lo[q] < hi[q]) {
                    const int md = (lo[q] + hi[q]) >> 1;
                    if (cps[md] <= ts[q]) lo[q] = md + 1; else hi[q] = md;
                }
                if (lo2[q] < hi2[q]) {
                    const int md = (lo2[q] + hi2[q]) >> 1;
                    if (cps[md + 1] <= te[q]) lo2[q] = md + 1; else hi2[q] = md;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = i0 + 32 * q;
            if (i >= S) continue;
            const int idx_lo = min(max(lo[q] - 1, 0), Sp - 1);
            const int idx_hi = min(max(lo2[q], 0), Sp - 1);
            const float w_outer = cy1[sl_pi<C>(idx_hi + 1)] - cy1[sl_pi<C>(idx_lo)];
            const float wi = wf[i];
            const float d = fmaxf(wi - w_outer, 0.f);
            ray_loss += d * d / (wi + LOSS_EPS);
            const float g = -2.f * d / (wi + LOSS_EPS) * g_scale;
            if (g != 0.f) {
                atomicAdd(E + sl_pi<C>(idx_hi + 1), g);
                atomicAdd(E + sl_pi<C>(idx_lo), -g);
            }
        }
    }
    t_inter += ray_loss;
    __syncwarp();
    // d loss / d w_j = sum of E over the entries behind j
    float e[C], sl = 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) sl += j0 + k < Sp ? E[sl_pi<C>(j0 + k)] : 0.f, e[k] = sl;
    const float incl_f = nvo_warp_scan_incl(sl, lane);
    float excl_f = __shfl_up_sync(0xffffffffu, incl_f, 1);
    if (lane == 0) excl_f = 0.f;
    const float tot = __shfl_sync(0xffffffffu, incl_f, 31) + E[sl_pi<C>(Sp)];
    float depth_l = 0.f, dwv[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        float dwj = tot - (excl_f + e[k]);
        if (use_depth && j0 + k < Sp) {
            const float t = (s0[k] + s1[k]) / 2.f, len = s1[k] - s0[k];
            const float diff = t - D;
            const float ex = expf(-(diff * diff) / two_sigma);
            depth_l += -logf(v[k] + LOSS_EPS) * ex * len;
            if (dmask && ex != 0.f) dwj += g_depth * (-1.f / (v[k] + LOSS_EPS)) * ex * len;
        }
        dwv[k] = dwj;
    }
    float* dwp = p.dw[l] + r * Sp;
    if (C % 4 == 0 && j0 + C <= Sp && (reinterpret_cast<uintptr_t>(dwp) & 15) == 0) {
#pragma unroll
        for (int k = 0; k < C; k += 4) *reinterpret_cast<float4*>(dwp + j0 + k) = make_float4(dwv[k], dwv[k + 1], dwv[k + 2], dwv[k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < C; ++k)
            if (j0 + k < Sp) dwp[j0 + k] = dwv[k];
    }
    if (dmask) t_depth += depth_l;
    __syncwarp();
}

// C0 / C1 > 0: the lane-blocked level path above with C = ceil(S_l / 32) samples per lane (the host picks the instantiation); 0: the chunked path
template <int C0, int C1>
__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK) k_step_losses(const __grid_constant__ StepLossP p) {
    extern __shared__ float smf[];
    __shared__ float red[RAYS_PER_BLOCK][5];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // three warps per ray: role 0 / 1 = that proposal level's interlevel + depth terms, role 2 = the field level's distortion + depth terms,
    // colour and normal losses (the levels' gradient buffers are disjoint, the five sums meet in the block reduction)
    const int64_t gw = (int64_t)blockIdx.x * RAYS_PER_BLOCK + wid;
    const int64_t r = gw / 3;
    const int role = (int)(gw - 3 * r);
    const bool live = r < p.B;
    const int S = p.S[2];
    const int npmax = 2 * (max(p.S[0], p.S[1]) + 1);  // rows indexed by sample carry the lane-blocked path's padding (< 2x)
    // per warp: wf[S], cf[S+1], m[S], then cy1 / cps / E [npmax] each
    float* wf = smf + wid * (3 * S + 1 + 3 * npmax);
    float* cf = wf + S;
    float* mid = cf + S + 1;
    float* cy1 = mid + S;
    float* cps = cy1 + npmax;
    float* E = cps + npmax;
    float t_rgb = 0.f, t_inter = 0.f, t_dist = 0.f, t_depth = 0.f, t_normal = 0.f;
    if (live) {
        const float B_f = (float)p.B;
        const bool use_depth = p.depth_gt != nullptr;
        const float D = use_depth ? __ldg(p.depth_gt + r) * __ldg(p.dnorm + r) : 0.f;
        const bool dmask = D > 0.f;
        const float two_sigma = 2.f * p.sigma;
        const float g_depth = p.mults[3] / B_f;
        for (int i = lane; i < S; i += 32) wf[i] = __ldg(p.w[2] + r * S + i);
        for (int i = lane; i <= S; i += 32) cf[i] = __ldg(p.sdist[2] + r * (S + 1) + i);
        __syncwarp();
        // ---- proposal levels: interlevel (fwd + bwd) and depth (fwd + bwd) ------------------------------------------------------------
        if (C0 > 0 && role == 0) sl_level_fast<(C0 > 0 ? C0 : 1)>(p, 0, r, lane, wf, cf, cy1, cps, E, D, use_depth, dmask, two_sigma, g_depth, B_f, t_inter, t_depth);
        if (C1 > 0 && role == 1) sl_level_fast<(C1 > 0 ? C1 : 1)>(p, 1, r, lane, wf, cf, cy1, cps, E, D, use_depth, dmask, two_sigma, g_depth, B_f, t_inter, t_depth);
        for (int l = 0; l < 2; ++l) {
            if (l != role || (l == 0 ? C0 : C1) > 0) continue;
            const int Sp = p.S[l], np = Sp + 1;
            const float* wp = p.w[l] + r * Sp;
            double carry = 0.0;
            for (int c0 = 0; c0 < Sp; c0 += 32) {
                const int i = c0 + lane;
                const float v = i < Sp ? __ldg(wp + i) : 0.f;
                const double incl = nvo_warp_scan_incl((double)v, lane);
                if (i < Sp) cy1[i + 1] = (float)(carry + incl);
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) cy1[0] = 0.f;
            for (int i = lane; i < np; i += 32) {
                cps[i] = __ldg(p.sdist[l] + r * np + i);
                E[i] = 0.f;
            }
            __syncwarp();
            const float g_scale = p.mults[1] / (B_f * (float)S);
            float ray_loss = 0.f;
            for (int i = lane; i < S; i += 32) {
                const float t0s = cf[i], t0e = cf[i + 1];
                int lo = 0, hi = Sp;
                while (lo < hi) {
                    const int md = (lo + hi) >> 1;
                    if (cps[md] <= t0s) lo = md + 1; else hi = md;
                }
                const int idx_lo = min(max(lo - 1, 0), Sp - 1);
                lo = 0, hi = Sp;
                while (lo < hi) {
                    const int md = (lo + hi) >> 1;
                    if (cps[md + 1] <= t0e) lo = md + 1; else hi = md;
                }
                const int idx_hi = min(max(lo, 0), Sp - 1);
                const float w_outer = cy1[idx_hi + 1] - cy1[idx_lo];
                const float wi = wf[i];
                const float d = fmaxf(wi - w_outer, 0.f);
                ray_loss += d * d / (wi + LOSS_EPS);
                const float g = -2.f * d / (wi + LOSS_EPS) * g_scale;
                if (g != 0.f) {
                    atomicAdd(E + idx_hi + 1, g);
                    atomicAdd(E + idx_lo, -g);
                }
            }
            t_inter += ray_loss;
            __syncwarp();
            float tot = 0.f;
            for (int i = lane; i < np; i += 32) tot += E[i];
            tot = nvo_warp_sum(tot);
            float carry_f = 0.f, depth_l = 0.f;
            const float* st = p.starts[l] + r * p.stride[l];
            const float* en = p.ends[l] + r * p.stride[l];
            for (int c0 = 0; c0 < Sp; c0 += 32) {
                const int j = c0 + lane;
                const float v = j < Sp ? E[j] : 0.f;
                const float incl = nvo_warp_scan_incl(v, lane);
                if (j < Sp) {
                    float dwj = 0.f;
                    dwj += tot - (carry_f + incl);
                    if (use_depth) {
                        const float s0 = __ldg(st + j), s1 = __ldg(en + j);
                        const float t = (s0 + s1) / 2.f, len = s1 - s0;
                        const float diff = t - D;
                        const float e = expf(-(diff * diff) / two_sigma);
                        const float wj = __ldg(wp + j);
                        depth_l += -logf(wj + LOSS_EPS) * e * len;
                        if (dmask && e != 0.f) dwj += g_depth * (-1.f / (wj + LOSS_EPS)) * e * len;
                    }
                    p.dw[l][r * Sp + j] = dwj;
                }
                carry_f += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (dmask) t_depth += depth_l;
            __syncwarp();
        }
        // ---- final level: distortion (fwd + bwd) and depth (fwd + bwd) ----------------------------------------------------------------
        for (int i = lane; i < S; i += 32) mid[i] = (cf[i + 1] + cf[i]) / 2.f;
        __syncwarp();
        if (role == 2) {
            const float g_dist = p.mults[2] / B_f;
            const float* st = p.starts[2] + r * p.stride[2];
            const float* en = p.ends[2] + r * p.stride[2];
            float depth_l = 0.f;
            for (int i = lane; i < S; i += 32) {
                const float mi = mid[i], wi = wf[i];
                float inner = 0.f;
                for (int j = 0; j < S; ++j) inner += wf[j] * fabsf(mi - mid[j]);
                const float delta = cf[i + 1] - cf[i];
                t_dist += wi * inner + wi * wi * delta / 3.f;
                float dwi = 0.f;
                dwi += g_dist * (2.f * inner + 2.f * wi * delta / 3.f);
                if (use_depth) {
                    const float s0 = __ldg(st + i), s1 = __ldg(en + i);
                    const float t = (s0 + s1) / 2.f, len = s1 - s0;
                    const float diff = t - D;
                    const float e = expf(-(diff * diff) / two_sigma);
                    depth_l += -logf(wi + LOSS_EPS) * e * len;
                    if (dmask && e != 0.f) dwi += g_depth * (-1.f / (wi + LOSS_EPS)) * e * len;
                }
                p.dw[2][r * S + i] = dwi;
            }
            if (dmask) t_depth += depth_l;
        }
        // ---- colour MSE (lanes 0..2) and MonoSDF normal loss (lane 3) -----------------------------------------------------------------
        if (role == 2 && lane < 3) {
            const float inv_n = 1.f / (3.f * B_f);
            const float d = __ldg(p.rgb + 3 * r + lane) - __ldg(p.rgb_gt + 3 * r + lane);
            t_rgb = d * d * inv_n;
            p.d_rgb[3 * r + lane] = p.mults[0] * 2.f * d * inv_n;
        }
        if (role == 2 && lane == 3 && p.normals) {
            const float inv_B = 1.f / B_f;
            float pv[3], gv[3], np_ = 0.f, ng = 0.f;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                pv[a] = __ldg(p.normals + 3 * r + a);
                gv[a] = __ldg(p.normal_gt + 3 * r + a);
                np_ += pv[a] * pv[a];
                ng += gv[a] * gv[a];
            }
            np_ = fmaxf(sqrtf(np_), 1e-12f);
            ng = fmaxf(sqrtf(ng), 1e-12f);
            float dot = 0.f, l1 = 0.f, dp[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                pv[a] /= np_;
                gv[a] /= ng;
                dot += pv[a] * gv[a];
                const float d = pv[a] - gv[a];
                l1 += fabsf(d);
                dp[a] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) - gv[a];
            }
            t_normal = (l1 + 1.f - dot) * inv_B;
            if (p.d_normals) {
                const float pd = pv[0] * dp[0] + pv[1] * dp[1] + pv[2] * dp[2];
#pragma unroll
                for (int a = 0; a < 3; ++a) p.d_normals[3 * r + a] = p.mults[4] * inv_B * (dp[a] - pv[a] * pd) / np_;
            }
        }
        t_inter *= 1.f / (B_f * (float)S);
        t_dist *= 1.f / B_f;
        t_depth *= 1.f / B_f;
    }
    t_rgb = nvo_warp_sum(t_rgb), t_inter = nvo_warp_sum(t_inter), t_dist = nvo_warp_sum(t_dist), t_depth = nvo_warp_sum(t_depth), t_normal = nvo_warp_sum(t_normal);
    if (lane == 0) red[wid][0] = t_rgb, red[wid][1] = t_inter, red[wid][2] = t_dist, red[wid][3] = t_depth, red[wid][4] = t_normal;
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x < 5) {
        float s = 0.f;
        for (int i = 0; i < RAYS_PER_BLOCK; ++i) s += red[i][threadIdx.x];
        atomicAdd(p.terms + threadIdx.x, s);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        float tot = 0.f;
        for (int i = 0; i < 5; ++i) tot += p.mults[i] * __ldcg(p.terms + i);
        *p.total = tot;
        *p.ticket = 0u;
    }
}

extern "C" int nvo_step_losses(void* stream, int64_t B, int32_t S0, int32_t S1, int32_t S2, const float* w0, const float* w1, const float* w2,
                               const float* sdist0, const float* sdist1, const float* sdist2, const float* starts0, const float* ends0, int64_t stride0,
                               const float* starts1, const float* ends1, int64_t stride1, const float* starts2, const float* ends2, int64_t stride2,
                               const float* rgb, const float* rgb_gt, const float* normals, const float* normal_gt, const float* depth_gt,
                               const float* directions_norm, float sigma, float mult_rgb, float mult_interlevel, float mult_distortion, float mult_depth,
                               float mult_normal, float* terms, float* total, void* ticket, float* dw0, float* dw1, float* dw2, float* d_rgb,
                               float* d_normals) {
    NVO_CHECK(B >= 0 && S0 >= 1 && S1 >= 1 && S2 >= 1 && S0 <= MAX_S && S1 <= MAX_S && S2 <= MAX_S, "step_losses: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(w0 && w1 && w2 && sdist0 && sdist1 && sdist2 && rgb && rgb_gt && terms && total && ticket && dw0 && dw1 && dw2 && d_rgb, "step_losses: null pointer");
    NVO_CHECK(!depth_gt || (directions_norm && starts0 && ends0 && starts1 && ends1 && starts2 && ends2), "step_losses: depth loss needs the sample intervals");
    NVO_CHECK(!normals || normal_gt, "step_losses: normal target missing");
    StepLossP p;
    p.B = B;
    p.S[0] = S0, p.S[1] = S1, p.S[2] = S2;
    p.w[0] = w0, p.w[1] = w1, p.w[2] = w2;
    p.sdist[0] = sdist0, p.sdist[1] = sdist1, p.sdist[2] = sdist2;
    p.starts[0] = starts0, p.starts[1] = starts1, p.starts[2] = starts2;
    p.ends[0] = ends0, p.ends[1] = ends1, p.ends[2] = ends2;
    p.stride[0] = stride0, p.stride[1] = stride1, p.stride[2] = stride2;
    p.rgb = rgb, p.rgb_gt = rgb_gt, p.normals = normals, p.normal_gt = normal_gt, p.depth_gt = depth_gt, p.dnorm = directions_norm, p.sigma = sigma;
    p.mults[0] = mult_rgb, p.mults[1] = mult_interlevel, p.mults[2] = mult_distortion, p.mults[3] = mult_depth, p.mults[4] = mult_normal;
    p.terms = terms, p.total = total, p.ticket = (unsigned int*)ticket;
    p.dw[0] = dw0, p.dw[1] = dw1, p.dw[2] = dw2, p.d_rgb = d_rgb, p.d_normals = d_normals;
    const size_t smem = sizeof(float) * RAYS_PER_BLOCK * (3 * (size_t)S2 + 1 + 3 * 2 * (size_t)(max(S0, S1) + 1));
    const int c0 = (S0 + 31) / 32, c1 = (S1 + 31) / 32;
    static const int fast = nvo_env_int("NVO_LOSS_FAST", 1);
    if (fast && c0 == 8 && c1 == 3) {  // NeRF-VO's 256 / 96 proposal samples
        NVO_CHECK(smem <= 48 * 1024, "step_losses: %zu bytes of shared memory per CTA exceed 48 KB", smem);
        k_step_losses<8, 3><<<ray_blocks(3 * B), 32 * RAYS_PER_BLOCK, smem, (cudaStream_t)stream>>>(p);
    } else if (fast && c0 == 2 && c1 == 1) {  // 64 / 32 samples: the small shapes of the parity tests walk the same code
        NVO_CHECK(smem <= 48 * 1024, "step_losses: %zu bytes of shared memory per CTA exceed 48 KB", smem);
        k_step_losses<2, 1><<<ray_blocks(3 * B), 32 * RAYS_PER_BLOCK, smem, (cudaStream_t)stream>>>(p);
    } else {
        if (smem > 48 * 1024) {
            NVO_CHECK(smem <= 200 * 1024, "step_losses: %zu bytes of shared memory per CTA exceed 200 KB", smem);
            cudaFuncSetAttribute(k_step_losses<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        k_step_losses<0, 0><<<ray_blocks(3 * B), 32 * RAYS_PER_BLOCK, smem, (cudaStream_t)stream>>>(p);
    }
    NVO_CUDA_LAUNCH_CHECK("step_losses");
    return 0;
}
