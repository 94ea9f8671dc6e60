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
