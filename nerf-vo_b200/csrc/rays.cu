// Per-ray operators: initial sampler, positions, alpha-compositing weights (+backward), PDF resampling,
// renderers (+backward).  One warp per ray; per-ray data staged in shared memory; prefix sums are warp scans
// carried in fp64 so that, like the reference's CPU cumsum (fp64 accumulator, fp32 result), every prefix is the
// correctly rounded fp32 value — that keeps searchsorted / median indices bit-exact against the oracle.
// Index-critical fp32 arithmetic uses __f*_rn intrinsics so the compiler cannot contract it into FMAs.
#include <math_constants.h>
#include "nvo_common.cuh"

#define RAYS_PER_BLOCK 4
#define MAX_S 1024

// ---- spacing functions of UniformLinDispPiecewiseSampler (ray_samplers.py:244-245) -------------------------
__device__ __forceinline__ float spacing_fn(float t) { return t < 1.f ? __fmul_rn(t, 0.5f) : __fsub_rn(1.f, __fdiv_rn(1.f, __fmul_rn(2.f, t))); }
__device__ __forceinline__ float spacing_fn_inv(float s) {
    return s < 0.5f ? __fmul_rn(2.f, s) : __fdiv_rn(1.f, __fsub_rn(2.f, __fmul_rn(2.f, s)));
}
// spacing_to_euclidean_fn (ray_samplers.py:114-115): inv(x*s_far + (1-x)*s_near)
__device__ __forceinline__ float to_euclid(float x, float s_near, float s_far) {
    return spacing_fn_inv(__fadd_rn(__fmul_rn(x, s_far), __fmul_rn(__fsub_rn(1.f, x), s_near)));
}

__global__ void k_sample_uniform(int64_t B, int S, const float* __restrict__ base, const float* __restrict__ jitter, const float* __restrict__ nears,
                                 const float* __restrict__ fars, float* __restrict__ sdist, float* __restrict__ ebins) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = S + 1;
    if (t >= B * n) return;
    const int64_t r = t / n;
    const int k = (int)(t - r * n);
    float b = __ldg(base + k);
    if (jitter) {
        // ray_samplers.py:105-109: lower + (upper - lower) * t_rand with centres (b[k+1]+b[k])/2
        const float lo = k == 0 ? b : __fmul_rn(__fadd_rn(b, __ldg(base + k - 1)), 0.5f);
        const float hi = k == S ? b : __fmul_rn(__fadd_rn(__ldg(base + k + 1), b), 0.5f);
        b = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), __ldg(jitter + r)));
    }
    sdist[t] = b;
    ebins[t] = to_euclid(b, spacing_fn(__ldg(nears + r)), spacing_fn(__ldg(fars + r)));
}

__global__ void k_sample_positions(int64_t B, int S, const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ starts,
                                   const float* __restrict__ ends, int64_t stride, float* __restrict__ pos) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * S) return;
    const int64_t r = t / S;
    const int k = (int)(t - r * S);
    const float se = __fadd_rn(__ldg(starts + r * stride + k), __ldg(ends + r * stride + k));
#pragma unroll
    for (int a = 0; a < 3; ++a)  // rays.py:55: origins + directions * (starts + ends) / 2
        pos[3 * t + a] = __fadd_rn(__ldg(o + 3 * r + a), __fmul_rn(__fmul_rn(__ldg(d + 3 * r + a), se), 0.5f));
}

__device__ __forceinline__ float nan_to_num(float v) {
    if (isnan(v)) return 0.f;
    if (isinf(v)) return v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    return v;
}

// ---- weights ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK) k_weights_fwd(int64_t B, int S, const float* __restrict__ starts, const float* __restrict__ ends, int64_t stride, const float* __restrict__ density,
                                                                     float* __restrict__ weights) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * RAYS_PER_BLOCK + (threadIdx.x >> 5);
    if (r >= B) return;
    const float* st = starts + r * stride;
    const float* en = ends + r * stride;
    double carry = 0.0;
    for (int c0 = 0; c0 < S; c0 += 32) {
        const int i = c0 + lane;
        float dd = 0.f;
        if (i < S) dd = __fmul_rn(__fsub_rn(__ldg(en + i), __ldg(st + i)), __ldg(density + r * S + i));
        // the reference's cumsum covers dd[0..S-2] only (rays.py:141); the last dd never enters a prefix
        const double incl = nvo_warp_scan_incl((double)dd, lane);
        const float excl = (float)(carry + incl - (double)dd);
        if (i < S) {
            const float alpha = __fsub_rn(1.f, expf(-dd));
            weights[r * S + i] = nan_to_num(__fmul_rn(alpha, expf(-excl)));
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
}

__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK) k_weights_bwd(int64_t B, int S, const float* __restrict__ starts, const float* __restrict__ ends, int64_t stride, const float* __restrict__ density,
                                                                     const float* __restrict__ dweights, float* __restrict__ ddensity) {
    // w_i = (1-e^{-a_i}) T_i, T_i = e^{-sum_{j<i} a_j}:  dL/da_i = g_i T_i e^{-a_i} - sum_{k>i} g_k w_k ; dL/dsigma_i = delta_i dL/da_i
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * RAYS_PER_BLOCK + (threadIdx.x >> 5);
    if (r >= B) return;
    const float* st = starts + r * stride;
    const float* en = ends + r * stride;
    // pass 1: total of g_k w_k
    double carry = 0.0, tot_gw = 0.0;
    for (int c0 = 0; c0 < S; c0 += 32) {
        const int i = c0 + lane;
        float dd = 0.f, g = 0.f;
        if (i < S) {
            dd = __fmul_rn(__fsub_rn(__ldg(en + i), __ldg(st + i)), __ldg(density + r * S + i));
            g = __ldg(dweights + r * S + i);
        }
        const double incl = nvo_warp_scan_incl((double)dd, lane);
        const float excl = (float)(carry + incl - (double)dd);
        float w = __fmul_rn(__fsub_rn(1.f, expf(-dd)), expf(-excl));
        if (!isfinite(w)) w = 0.f;  // nan_to_num zeroes the gradient path of non-finite weights
        tot_gw += nvo_warp_sum((double)(i < S ? g * w : 0.f));
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    carry = 0.0;
    double carry_gw = 0.0;
    for (int c0 = 0; c0 < S; c0 += 32) {
        const int i = c0 + lane;
        float dd = 0.f, g = 0.f, delta = 0.f;
        if (i < S) {
            delta = __fsub_rn(__ldg(en + i), __ldg(st + i));
            dd = __fmul_rn(delta, __ldg(density + r * S + i));
            g = __ldg(dweights + r * S + i);
        }
        const double incl = nvo_warp_scan_incl((double)dd, lane);
        const float excl = (float)(carry + incl - (double)dd);
        const float T = expf(-excl), e = expf(-dd);
        float w = __fmul_rn(__fsub_rn(1.f, e), T);
        const bool ok = isfinite(w);
        if (!ok) w = 0.f;
        const double gw = (i < S) ? (double)(g * w) : 0.0;
        const double incl_gw = nvo_warp_scan_incl(gw, lane);
        const double suffix = tot_gw - (carry_gw + incl_gw);  // sum_{k>i} g_k w_k
        if (i < S) {
            const float da = (ok ? g * T * e : 0.f) - (float)suffix;
            ddensity[r * S + i] = delta * da;
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
        carry_gw += __shfl_sync(0xffffffffu, incl_gw, 31);
    }
}

// ---- lane-blocked variants ------------------------------------------------------------------------------------
// Lane j owns the C CONSECUTIVE samples [C j, C j + C) of its ray (C = ceil(S / 32)): a prefix sum is then C serial fp64 adds, ONE warp scan of
// the lane totals and C adds, instead of one 5-step warp scan per 32-sample chunk (8 dependent shuffle chains at 256 samples).  Same fp64
// accumulation and per-entry fp32 rounding as the chunked kernels above, which stay as the path for other sample counts.
template <int C>
__device__ __forceinline__ void lane_block_scan(const double (&x)[C], int lane, double (&incl)[C], double& excl_lane, double& total) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < C; ++k) s += x[k], incl[k] = s;
    const double sc = nvo_warp_scan_incl(s, lane);
    excl_lane = __shfl_up_sync(0xffffffffu, sc, 1);
    if (lane == 0) excl_lane = 0.0;
    total = __shfl_sync(0xffffffffu, sc, 31);
}

template <int C>
__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK) k_weights_fwd_lb(int64_t B, int S, const float* __restrict__ starts, const float* __restrict__ ends,
                                                                        int64_t stride, const float* __restrict__ density, float* __restrict__ weights) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * RAYS_PER_BLOCK + (threadIdx.x >> 5);
    if (r >= B) return;
    const float* st = starts + r * stride;
    const float* en = ends + r * stride;
    const int j0 = C * lane;
    float dd[C];
    double x[C], incl[C], excl_lane, total;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int i = j0 + k;
        dd[k] = i < S ? __fmul_rn(__fsub_rn(__ldg(en + i), __ldg(st + i)), __ldg(density + r * S + i)) : 0.f;
        x[k] = (double)dd[k];
    }
    lane_block_scan<C>(x, lane, incl, excl_lane, total);
    float w[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const float excl = (float)(excl_lane + incl[k] - x[k]);
        w[k] = nan_to_num(__fmul_rn(__fsub_rn(1.f, expf(-dd[k])), expf(-excl)));
    }
    float* wp = weights + r * S;
    if (C % 4 == 0 && j0 + C <= S && (reinterpret_cast<uintptr_t>(wp) & 15) == 0) {
#pragma unroll
        for (int k = 0; k < C; k += 4) *reinterpret_cast<float4*>(wp + j0 + k) = make_float4(w[k], w[k + 1], w[k + 2], w[k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < C; ++k)
            if (j0 + k < S) wp[j0 + k] = w[k];
    }
}

template <int C>
__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK) k_weights_bwd_lb(int64_t B, int S, const float* __restrict__ starts, const float* __restrict__ ends,
                                                                        int64_t stride, const float* __restrict__ density, const float* __restrict__ dweights,
                                                                        float* __restrict__ ddensity) {
    // w_i = (1-e^{-a_i}) T_i, T_i = e^{-sum_{j<i} a_j}:  dL/da_i = g_i T_i e^{-a_i} - sum_{k>i} g_k w_k ; dL/dsigma_i = delta_i dL/da_i
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * RAYS_PER_BLOCK + (threadIdx.x >> 5);
    if (r >= B) return;
    const float* st = starts + r * stride;
    const float* en = ends + r * stride;
    const int j0 = C * lane;
    float delta[C], g[C], dd[C];
    double x[C], incl[C], excl_lane, total;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int i = j0 + k;
        const bool in = i < S;
        delta[k] = in ? __fsub_rn(__ldg(en + i), __ldg(st + i)) : 0.f;
        dd[k] = in ? __fmul_rn(delta[k], __ldg(density + r * S + i)) : 0.f;
        g[k] = in ? __ldg(dweights + r * S + i) : 0.f;
        x[k] = (double)dd[k];
    }
    lane_block_scan<C>(x, lane, incl, excl_lane, total);
    float first[C];  // g_i T_i e^{-a_i} where the weight is finite
    double gw[C], incl_gw[C], excl_gw, tot_gw;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const float excl = (float)(excl_lane + incl[k] - x[k]);
        const float T = expf(-excl), e = expf(-dd[k]);
        float w = __fmul_rn(__fsub_rn(1.f, e), T);
        const bool ok = isfinite(w);
        if (!ok) w = 0.f;  // nan_to_num zeroes the gradient path of non-finite weights
        gw[k] = j0 + k < S ? (double)(g[k] * w) : 0.0;
        first[k] = ok ? g[k] * T * e : 0.f;
    }
    lane_block_scan<C>(gw, lane, incl_gw, excl_gw, tot_gw);
    float out[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const double suffix = tot_gw - (excl_gw + incl_gw[k]);  // sum_{k' > i} g_k' w_k'
        out[k] = delta[k] * (first[k] - (float)suffix);
    }
    float* dp = ddensity + r * S;
    if (C % 4 == 0 && j0 + C <= S && (reinterpret_cast<uintptr_t>(dp) & 15) == 0) {
#pragma unroll
        for (int k = 0; k < C; k += 4) *reinterpret_cast<float4*>(dp + j0 + k) = make_float4(out[k], out[k + 1], out[k + 2], out[k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < C; ++k)
            if (j0 + k < S) dp[j0 + k] = out[k];
    }
}

// ---- PDF resampling -------------------------------------------------------------------------------------------
// smem per warp: cdf[S_in+1], bins[S_in+1]
template <int C>
__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK) k_pdf_resample(int64_t B, int S_in, int S_out, const float* __restrict__ weights,
                                                                      const float* __restrict__ sdist_in, const float* __restrict__ u_base,
                                                                      const float* __restrict__ jitter, float anneal, const float* __restrict__ anneal_dev, float pad,
                                                                      const float* __restrict__ nears,
                                                                      const float* __restrict__ fars, float* __restrict__ sdist_out,
                                                                      float* __restrict__ ebins_out, int32_t* __restrict__ inds_out) {
    extern __shared__ float smf[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * RAYS_PER_BLOCK + wid;
    if (r >= B) return;
    const int n_in = S_in + 1, n_out = S_out + 1;
    float* cdf = smf + wid * 2 * n_in;
    float* bins = cdf + n_in;
    const float eps = 1e-5f;
    if (anneal_dev) anneal = __ldg(anneal_dev);  // device scalar: follows the anneal schedule across CUDA-graph replays
    // pass 1: w = pow(w, anneal) + padding ; sum (ray_samplers.py:305-308, :602)
    double sum_d = 0.0;
    for (int i = lane; i < S_in; i += 32) {
        float w = __ldg(weights + r * S_in + i);
        if (anneal != 1.f) w = powf(w, anneal);
        w = __fadd_rn(w, pad);
        cdf[i + 1] = w;  // stash
        sum_d += (double)w;
    }
    for (int i = lane; i < n_in; i += 32) bins[i] = __ldg(sdist_in + r * n_in + i);
    float w_sum = (float)nvo_warp_sum(sum_d);
    const float padding = fmaxf(__fsub_rn(eps, w_sum), 0.f);
    const float pad_each = __fdiv_rn(padding, (float)S_in);
    w_sum = __fadd_rn(w_sum, padding);
    __syncwarp();
    // pass 2: pdf = w / sum ; cdf = min(1, cumsum(pdf)) ; cdf = [0, cdf]
    if (C > 0) {  // lane-blocked cumsum (C = ceil(S_in / 32) samples per lane), see lane_block_scan
        constexpr int CC = C > 0 ? C : 1;
        double x[CC], incl[CC], excl_lane, total;
#pragma unroll
        for (int k = 0; k < CC; ++k) {
            const int i = CC * lane + k;
            x[k] = i < S_in ? (double)__fdiv_rn(__fadd_rn(cdf[i + 1], pad_each), w_sum) : 0.0;
        }
        __syncwarp();
        lane_block_scan<CC>(x, lane, incl, excl_lane, total);
#pragma unroll
        for (int k = 0; k < CC; ++k) {
            const int i = CC * lane + k;
            if (i < S_in) cdf[i + 1] = fminf(1.f, (float)(excl_lane + incl[k]));
        }
    } else {
        double carry = 0.0;
        for (int c0 = 0; c0 < S_in; c0 += 32) {
            const int i = c0 + lane;
            float pdf = 0.f;
            if (i < S_in) pdf = __fdiv_rn(__fadd_rn(cdf[i + 1], pad_each), w_sum);
            const double incl = nvo_warp_scan_incl((double)pdf, lane);
            if (i < S_in) cdf[i + 1] = fminf(1.f, (float)(carry + incl));
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    if (lane == 0) cdf[0] = 0.f;
    __syncwarp();
    const float s_near = spacing_fn(__ldg(nears + r)), s_far = spacing_fn(__ldg(fars + r));
    const float jit = jitter ? __fdiv_rn(__ldg(jitter + r), (float)n_out) : 0.f;
    // four output samples per lane and trip: their binary searches advance together (independent chains hide the shared-memory latency)
    int iters = 0;
    while ((1 << iters) < n_in + 1) ++iters;
    for (int k0 = lane; k0 < n_out; k0 += 128) {
        float u[4];
        int lo[4], hi[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + 32 * q;
            const bool act = k < n_out;
            u[q] = act ? (jitter ? __fadd_rn(__ldg(u_base + k), jit) : __ldg(u_base + k)) : 0.f;
            lo[q] = 0, hi[q] = act ? n_in : 0;
        }
        // searchsorted(cdf, u, side="right"): number of entries <= u
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (lo[q] < hi[q]) {
                    const int mid = (lo[q] + hi[q]) >> 1;
                    if (cdf[mid] <= u[q])
                        lo[q] = mid + 1;
                    else
                        hi[q] = mid;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + 32 * q;
            if (k >= n_out) continue;
            const int ind = lo[q];
            if (inds_out) inds_out[r * n_out + k] = ind;
            const int below = min(max(ind - 1, 0), S_in), above = min(max(ind, 0), S_in);
            const float c0 = cdf[below], c1 = cdf[above], b0 = bins[below], b1 = bins[above];
            float t = __fdiv_rn(__fsub_rn(u[q], c0), __fsub_rn(c1, c0));
            t = nan_to_num(t);
            t = fminf(fmaxf(t, 0.f), 1.f);
            const float b = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
            sdist_out[r * n_out + k] = b;
            ebins_out[r * n_out + k] = to_euclid(b, s_near, s_far);
        }
    }
}

// ---- renderers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_pos(float* addr, float v) { atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v)); }
__device__ __forceinline__ void atomic_max_pos(float* addr, float v) { atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v)); }

__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK)
    k_render_fwd(int64_t B, int S, int eval_mode, const float* __restrict__ starts, const float* __restrict__ ends, int64_t stride, const float* __restrict__ weights, const float* __restrict__ rgb,
                 const float* __restrict__ normals, const float* __restrict__ pred_normals, float* __restrict__ out_rgb, float* __restrict__ out_acc,
                 float* __restrict__ out_dexp, float* __restrict__ minmax, float* __restrict__ out_dmed, int32_t* __restrict__ out_midx,
                 float* __restrict__ out_n, float* __restrict__ out_pn) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * RAYS_PER_BLOCK + (threadIdx.x >> 5);
    if (r >= B) return;
    const float* st = starts + r * stride;
    const float* en = ends + r * stride;
    const float* w_ = weights + r * S;
    float acc = 0.f, c[3] = {0, 0, 0}, wt = 0.f, n[3] = {0, 0, 0}, pn[3] = {0, 0, 0};
    float tmin = CUDART_INF_F, tmax = -CUDART_INF_F;
    int below_half = 0;  // searchsorted(cumsum(w), 0.5, left) = #{cw < 0.5}
    double carry = 0.0;
    for (int c0 = 0; c0 < S; c0 += 32) {
        const int i = c0 + lane;
        const bool live = i < S;
        const float w = live ? __ldg(w_ + i) : 0.f;
        const float t = live ? __fmul_rn(__fadd_rn(__ldg(st + i), __ldg(en + i)), 0.5f) : 0.f;
        if (live) {
            acc += w;
            wt += w * t;
            tmin = fminf(tmin, t);
            tmax = fmaxf(tmax, t);
            if (rgb) {
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    float v = __ldg(rgb + (r * S + i) * 3 + a);
                    if (eval_mode) v = nan_to_num(v);
                    c[a] += w * v;
                }
            }
            if (normals)
#pragma unroll
                for (int a = 0; a < 3; ++a) n[a] += w * __ldg(normals + (r * S + i) * 3 + a);
            if (pred_normals)
#pragma unroll
                for (int a = 0; a < 3; ++a) pn[a] += w * __ldg(pred_normals + (r * S + i) * 3 + a);
        }
        if (out_dmed || out_midx) {
            const double incl = nvo_warp_scan_incl((double)w, lane);
            const float cw = (float)(carry + incl);
            below_half += __popc(__ballot_sync(0xffffffffu, live && cw < 0.5f));
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    acc = nvo_warp_sum(acc);
    wt = nvo_warp_sum(wt);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        c[a] = nvo_warp_sum(c[a]);
        n[a] = nvo_warp_sum(n[a]);
        pn[a] = nvo_warp_sum(pn[a]);
    }
    tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, 16));
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 16));
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    }
    if (lane == 0) {
        if (out_rgb && rgb) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float last = __ldg(rgb + (r * S + S - 1) * 3 + a);
                if (eval_mode) last = nan_to_num(last);
                float v = c[a] + last * (1.f - acc);  // renderers.py:108-116 background 'last_sample'
                if (eval_mode) v = fminf(fmaxf(v, 0.f), 1.f);
                out_rgb[r * 3 + a] = v;
            }
        }
        if (out_acc) out_acc[r] = acc;
        if (out_dexp) {
            out_dexp[r] = wt / (acc + 1e-10f);
            if (minmax) {  // mid-steps are positive: int ordering == float ordering
                atomic_min_pos(minmax, tmin);
                atomic_max_pos(minmax + 1, tmax);
            }
        }
        if (out_dmed || out_midx) {
            const int idx = min(max(below_half, 0), S - 1);
            if (out_midx) out_midx[r] = idx;
            if (out_dmed) out_dmed[r] = __fmul_rn(__fadd_rn(__ldg(st + idx), __ldg(en + idx)), 0.5f);
        }
        if (out_n && normals) {
            const float inv = 1.f / (sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) + 1e-10f);
#pragma unroll
            for (int a = 0; a < 3; ++a) out_n[r * 3 + a] = (n[a] * inv + 1.f) / 2.f;
        }
        if (out_pn && pred_normals) {
            const float inv = 1.f / (sqrtf(pn[0] * pn[0] + pn[1] * pn[1] + pn[2] * pn[2]) + 1e-10f);
#pragma unroll
            for (int a = 0; a < 3; ++a) out_pn[r * 3 + a] = (pn[a] * inv + 1.f) / 2.f;
        }
    }
}

__global__ void k_init_minmax(float* minmax) {
    minmax[0] = CUDART_INF_F;
    minmax[1] = -CUDART_INF_F;
}

__global__ void k_clip_depth(int64_t B, const float* __restrict__ minmax, float* __restrict__ depth) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B) return;
    depth[t] = fminf(fmaxf(depth[t], minmax[0]), minmax[1]);
}

// gradient of out = (N/(|N|+eps) + 1)/2 w.r.t. N, given d_out
__device__ __forceinline__ void normalize_bwd(const float N[3], const float dout[3], float dN[3]) {
    const float r = sqrtf(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
    const float re = r + 1e-10f;
    const float dot = N[0] * dout[0] + N[1] * dout[1] + N[2] * dout[2];
    const float k = r > 0.f ? dot / (r * re * re) : 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) dN[a] = 0.5f * (dout[a] / re - N[a] * k);
}

__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK)
    k_render_bwd(int64_t B, int S, const float* __restrict__ starts, const float* __restrict__ ends, int64_t stride, const float* __restrict__ weights, const float* __restrict__ rgb,
                 const float* __restrict__ normals, const float* __restrict__ pred_normals, const float* __restrict__ d_rgb, const float* __restrict__ d_acc,
                 const float* __restrict__ d_dexp, const float* __restrict__ minmax, const float* __restrict__ d_n, const float* __restrict__ d_pn,
                 int accumulate, float* __restrict__ dweights, float* __restrict__ drgb, float* __restrict__ dpn) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * RAYS_PER_BLOCK + (threadIdx.x >> 5);
    if (r >= B) return;
    const float* st = starts + r * stride;
    const float* en = ends + r * stride;
    const float* w_ = weights + r * S;
    // pass 1: per-ray reductions the gradients need
    float acc = 0.f, wt = 0.f, N[3] = {0, 0, 0}, PN[3] = {0, 0, 0};
    for (int i = lane; i < S; i += 32) {
        const float w = __ldg(w_ + i);
        acc += w;
        if (d_dexp) wt += w * __fmul_rn(__fadd_rn(__ldg(st + i), __ldg(en + i)), 0.5f);
        if (d_n && normals)
#pragma unroll
            for (int a = 0; a < 3; ++a) N[a] += w * __ldg(normals + (r * S + i) * 3 + a);
        if (d_pn && pred_normals)
#pragma unroll
            for (int a = 0; a < 3; ++a) PN[a] += w * __ldg(pred_normals + (r * S + i) * 3 + a);
    }
    acc = nvo_warp_sum(acc);
    wt = nvo_warp_sum(wt);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        N[a] = nvo_warp_sum(N[a]);
        PN[a] = nvo_warp_sum(PN[a]);
    }
    float g_rgb[3] = {0, 0, 0}, last[3] = {0, 0, 0}, dN[3] = {0, 0, 0}, dPN[3] = {0, 0, 0};
    if (d_rgb && rgb)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            g_rgb[a] = __ldg(d_rgb + r * 3 + a);
            last[a] = __ldg(rgb + (r * S + S - 1) * 3 + a);
        }
    if (d_n && normals) {
        const float g[3] = {__ldg(d_n + r * 3), __ldg(d_n + r * 3 + 1), __ldg(d_n + r * 3 + 2)};
        normalize_bwd(N, g, dN);
    }
    if (d_pn && pred_normals) {
        const float g[3] = {__ldg(d_pn + r * 3), __ldg(d_pn + r * 3 + 1), __ldg(d_pn + r * 3 + 2)};
        normalize_bwd(PN, g, dPN);
    }
    const float g_acc = d_acc ? __ldg(d_acc + r) : 0.f;
    float g_dexp = 0.f, dexp = 0.f;
    if (d_dexp) {
        dexp = wt / (acc + 1e-10f);
        const bool inside = !minmax || (dexp >= minmax[0] && dexp <= minmax[1]);  // clip passes gradient inside the range
        g_dexp = inside ? __ldg(d_dexp + r) : 0.f;
    }
    const float g_last_dot = g_rgb[0] * last[0] + g_rgb[1] * last[1] + g_rgb[2] * last[2];
    for (int i = lane; i < S; i += 32) {
        const float w = __ldg(w_ + i);
        float dw = g_acc - g_last_dot;  // d(acc) and d(last*(1-acc))
        if (d_rgb && rgb) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                dw += g_rgb[a] * __ldg(rgb + (r * S + i) * 3 + a);
                if (drgb) drgb[(r * S + i) * 3 + a] = g_rgb[a] * (w + (i == S - 1 ? 1.f - acc : 0.f));
            }
        } else if (drgb) {
#pragma unroll
            for (int a = 0; a < 3; ++a) drgb[(r * S + i) * 3 + a] = 0.f;
        }
        if (d_dexp) {
            const float t = __fmul_rn(__fadd_rn(__ldg(st + i), __ldg(en + i)), 0.5f);
            dw += g_dexp * (t - dexp) / (acc + 1e-10f);
        }
        if (d_n && normals)
#pragma unroll
            for (int a = 0; a < 3; ++a) dw += dN[a] * __ldg(normals + (r * S + i) * 3 + a);
        if (pred_normals && (d_pn || dpn)) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (d_pn) dw += dPN[a] * __ldg(pred_normals + (r * S + i) * 3 + a);
                if (dpn) dpn[(r * S + i) * 3 + a] = d_pn ? w * dPN[a] : 0.f;
            }
        }
        if (accumulate)
            dweights[r * S + i] += dw;
        else
            dweights[r * S + i] = dw;
    }
}

// ---- C ABI -----------------------------------------------------------------------------------------------------
static inline unsigned int ray_blocks(int64_t B) { return (unsigned int)((B + RAYS_PER_BLOCK - 1) / RAYS_PER_BLOCK); }

extern "C" int nvo_sample_uniform(void* stream, int64_t B, int32_t S, const float* base_bins, const float* jitter, const float* nears,
                                  const float* fars, float* sdist, float* ebins) {
    NVO_CHECK(B >= 0 && S >= 1 && S <= MAX_S, "sample_uniform: bad shape B=%lld S=%d", (long long)B, S);
    if (B == 0) return 0;
    NVO_CHECK(base_bins && nears && fars && sdist && ebins, "sample_uniform: null pointer");
    k_sample_uniform<<<nvo_blocks(B * (S + 1), 256), 256, 0, (cudaStream_t)stream>>>(B, S, base_bins, jitter, nears, fars, sdist, ebins);
    NVO_CUDA_LAUNCH_CHECK("sample_uniform");
    return 0;
}

extern "C" int nvo_sample_positions(void* stream, int64_t B, int32_t S, const float* origins, const float* directions, const float* starts,
                                    const float* ends, int64_t stride, float* pos) {
    NVO_CHECK(B >= 0 && S >= 1, "sample_positions: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(origins && directions && starts && ends && pos, "sample_positions: null pointer");
    k_sample_positions<<<nvo_blocks(B * S, 256), 256, 0, (cudaStream_t)stream>>>(B, S, origins, directions, starts, ends, stride, pos);
    NVO_CUDA_LAUNCH_CHECK("sample_positions");
    return 0;
}

extern "C" int nvo_weights_forward(void* stream, int64_t B, int32_t S, const float* starts, const float* ends, int64_t stride, const float* density,
                                   float* weights) {
    NVO_CHECK(B >= 0 && S >= 1, "weights_forward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(starts && ends && density && weights, "weights_forward: null pointer");
    static const int lb = nvo_env_int("NVO_RAYS_LANE_BLOCKED", 1);
    const int c = lb ? (S + 31) / 32 : 0;
    if (c == 8)
        k_weights_fwd_lb<8><<<ray_blocks(B), 32 * RAYS_PER_BLOCK, 0, (cudaStream_t)stream>>>(B, S, starts, ends, stride, density, weights);
    else if (c == 3)
        k_weights_fwd_lb<3><<<ray_blocks(B), 32 * RAYS_PER_BLOCK, 0, (cudaStream_t)stream>>>(B, S, starts, ends, stride, density, weights);
    else
        k_weights_fwd<<<ray_blocks(B), 32 * RAYS_PER_BLOCK, 0, (cudaStream_t)stream>>>(B, S, starts, ends, stride, density, weights);
    NVO_CUDA_LAUNCH_CHECK("weights_forward");
    return 0;
}

extern "C" int nvo_weights_backward(void* stream, int64_t B, int32_t S, const float* starts, const float* ends, int64_t stride, const float* density,
                                    const float* dweights, float* ddensity) {
    NVO_CHECK(B >= 0 && S >= 1, "weights_backward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(starts && ends && density && dweights && ddensity, "weights_backward: null pointer");
    static const int lb = nvo_env_int("NVO_RAYS_LANE_BLOCKED", 1);
    const int c = lb ? (S + 31) / 32 : 0;
    if (c == 8)
        k_weights_bwd_lb<8><<<ray_blocks(B), 32 * RAYS_PER_BLOCK, 0, (cudaStream_t)stream>>>(B, S, starts, ends, stride, density, dweights, ddensity);
    else if (c == 3)
        k_weights_bwd_lb<3><<<ray_blocks(B), 32 * RAYS_PER_BLOCK, 0, (cudaStream_t)stream>>>(B, S, starts, ends, stride, density, dweights, ddensity);
    else
        k_weights_bwd<<<ray_blocks(B), 32 * RAYS_PER_BLOCK, 0, (cudaStream_t)stream>>>(B, S, starts, ends, stride, density, dweights, ddensity);
    NVO_CUDA_LAUNCH_CHECK("weights_backward");
    return 0;
}

extern "C" int nvo_pdf_resample(void* stream, int64_t B, int32_t S_in, int32_t S_out, const float* weights, const float* sdist_in, const float* u_base,
                                const float* jitter, float anneal, const float* anneal_dev, float histogram_padding, const float* nears, const float* fars,
                                float* sdist_out, float* ebins_out, int32_t* inds) {
    NVO_CHECK(B >= 0 && S_in >= 1 && S_in <= MAX_S && S_out >= 1 && S_out <= MAX_S, "pdf_resample: bad shape B=%lld S_in=%d S_out=%d", (long long)B, S_in, S_out);
    if (B == 0) return 0;
    NVO_CHECK(weights && sdist_in && u_base && nears && fars && sdist_out && ebins_out, "pdf_resample: null pointer");
    const size_t smem = sizeof(float) * 2 * (S_in + 1) * RAYS_PER_BLOCK;
    static const int lb = nvo_env_int("NVO_RAYS_LANE_BLOCKED", 1);
    const int c = lb ? (S_in + 31) / 32 : 0;
    if (c == 8)
        k_pdf_resample<8><<<ray_blocks(B), 32 * RAYS_PER_BLOCK, smem, (cudaStream_t)stream>>>(B, S_in, S_out, weights, sdist_in, u_base, jitter, anneal, anneal_dev,
                                                                                              histogram_padding, nears, fars, sdist_out, ebins_out, inds);
    else if (c == 3)
        k_pdf_resample<3><<<ray_blocks(B), 32 * RAYS_PER_BLOCK, smem, (cudaStream_t)stream>>>(B, S_in, S_out, weights, sdist_in, u_base, jitter, anneal, anneal_dev,
                                                                                              histogram_padding, nears, fars, sdist_out, ebins_out, inds);
    else
        k_pdf_resample<0><<<ray_blocks(B), 32 * RAYS_PER_BLOCK, smem, (cudaStream_t)stream>>>(B, S_in, S_out, weights, sdist_in, u_base, jitter, anneal, anneal_dev,
                                                                                              histogram_padding, nears, fars, sdist_out, ebins_out, inds);
    NVO_CUDA_LAUNCH_CHECK("pdf_resample");
    return 0;
}

extern "C" int nvo_render_forward(void* stream, int64_t B, int32_t S, int32_t eval_mode, const float* starts, const float* ends, int64_t stride,
                                  const float* weights, const float* rgb,
                                  const float* normals, const float* pred_normals, float* out_rgb, float* out_acc, float* out_depth_expected,
                                  float* minmax, float* out_depth_median, int32_t* out_median_idx, float* out_normals, float* out_pred_normals) {
    NVO_CHECK(B >= 0 && S >= 1, "render_forward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(starts && ends && weights, "render_forward: null pointer");
    if (minmax && out_depth_expected) k_init_minmax<<<1, 1, 0, (cudaStream_t)stream>>>(minmax);
    k_render_fwd<<<ray_blocks(B), 32 * RAYS_PER_BLOCK, 0, (cudaStream_t)stream>>>(B, S, eval_mode, starts, ends, stride, weights, rgb, normals, pred_normals, out_rgb, out_acc,
                                                                                  out_depth_expected, minmax, out_depth_median, out_median_idx,
                                                                                  out_normals, out_pred_normals);
    NVO_CUDA_LAUNCH_CHECK("render_forward");
    return 0;
}

extern "C" int nvo_clip_depth(void* stream, int64_t B, const float* minmax, float* depth) {
    NVO_CHECK(B >= 0, "clip_depth: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(minmax && depth, "clip_depth: null pointer");
    k_clip_depth<<<nvo_blocks(B, 256), 256, 0, (cudaStream_t)stream>>>(B, minmax, depth);
    NVO_CUDA_LAUNCH_CHECK("clip_depth");
    return 0;
}

extern "C" int nvo_render_backward(void* stream, int64_t B, int32_t S, const float* starts, const float* ends, int64_t stride, const float* weights,
                                   const float* rgb, const float* normals,
                                   const float* pred_normals, const float* d_out_rgb, const float* d_out_acc, const float* d_out_depth_expected,
                                   const float* minmax, const float* d_out_normals, const float* d_out_pred_normals, int32_t accumulate_dweights,
                                   float* dweights, float* drgb, float* dpred_normals) {
    NVO_CHECK(B >= 0 && S >= 1, "render_backward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(starts && ends && weights && dweights, "render_backward: null pointer");
    k_render_bwd<<<ray_blocks(B), 32 * RAYS_PER_BLOCK, 0, (cudaStream_t)stream>>>(B, S, starts, ends, stride, weights, rgb, normals, pred_normals, d_out_rgb, d_out_acc,
                                                                                  d_out_depth_expected, minmax, d_out_normals, d_out_pred_normals,
                                                                                  accumulate_dweights, dweights, drgb, dpred_normals);
    NVO_CUDA_LAUNCH_CHECK("render_backward");
    return 0;
}

// ================================================================================================================================
// d loss / d (ray origin, ray direction) from d loss / d x, x = the normalised contracted sample positions the hash grids read:
// x = selector * (contract_Linf(o + d (start + end) / 2) + 2) / 4  (rays.py:49-58, spatial_distortions.py:67-69, nerfacto_field.py:204-209).
// One warp per ray: every lane back-propagates its samples through the selector mask, the normalisation and the contraction Jacobian, the
// warp sums dL/dp (origin) and t dL/dp (direction) and adds them to the ray's rows with six atomics (the three sampling levels and the
// field's direction encoding accumulate into the same buffers, possibly from different streams).
// ================================================================================================================================
__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK) k_position_backward(int64_t B, int S, const float* __restrict__ o, const float* __restrict__ d,
                                                                           const float* __restrict__ starts, const float* __restrict__ ends, int64_t stride,
                                                                           const float* __restrict__ dx, float* __restrict__ d_o, float* __restrict__ d_d) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * RAYS_PER_BLOCK + wid;
    if (r >= B) return;
    const float ox = __ldg(o + 3 * r), oy = __ldg(o + 3 * r + 1), oz = __ldg(o + 3 * r + 2);
    const float dxr = __ldg(d + 3 * r), dyr = __ldg(d + 3 * r + 1), dzr = __ldg(d + 3 * r + 2);
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < S; k += 32) {
        const float se = __fadd_rn(__ldg(starts + r * stride + k), __ldg(ends + r * stride + k));
        const float t = se * 0.5f;
        const float p[3] = {__fadd_rn(ox, __fmul_rn(__fmul_rn(dxr, se), 0.5f)), __fadd_rn(oy, __fmul_rn(__fmul_rn(dyr, se), 0.5f)),
                            __fadd_rn(oz, __fmul_rn(__fmul_rn(dzr, se), 0.5f))};
        const float a[3] = {fabsf(p[0]), fabsf(p[1]), fabsf(p[2])};
        const float mag = fmaxf(a[0], fmaxf(a[1], a[2]));
        const bool inside = mag < 1.f;
        // forward values again (selector)
        bool sel = true;
        float sfac = 1.f;
        if (!inside) sfac = (2.f - 1.f / mag) / mag;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float q = __fmul_rn(__fadd_rn(inside ? p[i] : __fmul_rn(__fsub_rn(2.f, __fdiv_rn(1.f, mag)), __fdiv_rn(p[i], mag)), 2.f), 0.25f);
            sel = sel && (q > 0.f) && (q < 1.f);
        }
        if (!sel) continue;
        const int64_t smp = r * S + k;
        const float g[3] = {0.25f * __ldg(dx + 3 * smp), 0.25f * __ldg(dx + 3 * smp + 1), 0.25f * __ldg(dx + 3 * smp + 2)};  // d loss / d contracted
        float dp[3];
        if (inside) {
            dp[0] = g[0], dp[1] = g[1], dp[2] = g[2];
        } else {
            // c = s(m) p with s = 2/m - 1/m^2, m = |p_k| the largest component: dL/dp_j = s g_j + [j == k] sign(p_k) s'(m) (g . p)
            const int kmax = (a[0] >= a[1] && a[0] >= a[2]) ? 0 : (a[1] >= a[2] ? 1 : 2);
            const float ds = -2.f / (mag * mag) + 2.f / (mag * mag * mag);
            const float gp = g[0] * p[0] + g[1] * p[1] + g[2] * p[2];
#pragma unroll
            for (int i = 0; i < 3; ++i) dp[i] = sfac * g[i] + (i == kmax ? (p[i] < 0.f ? -1.f : 1.f) * ds * gp : 0.f);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            acc[i] += dp[i];
            acc[3 + i] += t * dp[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) acc[i] = nvo_warp_sum(acc[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (d_o && acc[i] != 0.f) atomicAdd(d_o + 3 * r + i, acc[i]);
            if (d_d && acc[3 + i] != 0.f) atomicAdd(d_d + 3 * r + i, acc[3 + i]);
        }
    }
}

extern "C" int nvo_position_backward(void* stream, int64_t B, int32_t S, const float* origins, const float* directions, const float* starts,
                                     const float* ends, int64_t stride, const float* dx, float* d_origins, float* d_directions) {
    NVO_CHECK(B >= 0 && S >= 1, "position_backward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(origins && directions && starts && ends && dx && (d_origins || d_directions), "position_backward: null pointer");
    k_position_backward<<<ray_blocks(B), 32 * RAYS_PER_BLOCK, 0, (cudaStream_t)stream>>>(B, S, origins, directions, starts, ends, stride, dx, d_origins,
                                                                                          d_directions);
    NVO_CUDA_LAUNCH_CHECK("position_backward");
    return 0;
}
