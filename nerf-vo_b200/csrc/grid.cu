// Multiresolution hash grid: gather / trilinear forward, atomic-scatter backward, input gradient.
// Semantics = the reference's torch path (NS/field_components/encodings.py:405-465):
//   scaled = x * scale_l (fp32), corners floor/ceil, hash (x*1 ^ y*2654435761 ^ z*805459861) mod 2^k + l*2^k,
//   interpolation order x, y, z with separately-rounded mul/add (no FMA contraction) so that the fp32 path
//   reproduces the reference bit for bit.
// Work decomposition: one thread per (sample, level) with level fastest, so a warp's 32 threads cover
// 32/L consecutive samples: x loads are near-broadcast, the y / dy accesses of a warp are one contiguous
// 256-byte run (fp32) and every thread keeps 8 independent 8-byte (float2) or 4-byte (half2) gathers in flight.
#include "grid_common.cuh"
#include <stdlib.h>

__device__ __forceinline__ void store_feat(float* y, int64_t t, float a, float b) { reinterpret_cast<float2*>(y)[t] = make_float2(a, b); }
__device__ __forceinline__ void store_feat(__half* y, int64_t t, float a, float b) { reinterpret_cast<__half2*>(y)[t] = __floats2half2_rn(a, b); }
__device__ __forceinline__ float2 load_feat(const float* y, int64_t t) { return __ldg(reinterpret_cast<const float2*>(y) + t); }
__device__ __forceinline__ float2 load_feat(const __half* y, int64_t t) { return __half22float2(__ldg(reinterpret_cast<const __half2*>(y) + t)); }
// dy of (sample s, level l): row-major [n, 2L] (t = s*L + l) or, tmf != 0, fp32 tile-major [tile][2L][128] (mlp_tc.cu's dx)
template <typename OutT>
__device__ __forceinline__ float2 load_dy(const OutT* dy, int64_t s, int l, int L, int tmf) {
    return load_feat(dy, s * L + l);
}
template <>
__device__ __forceinline__ float2 load_dy<float>(const float* dy, int64_t s, int l, int L, int tmf) {
    if (!tmf) return load_feat(dy, s * L + l);
    const float* b = dy + (((s >> 7) * (2 * L) + 2 * l) << 7) + (s & 127);
    return make_float2(__ldg(b), __ldg(b + 128));
}

template <typename RowT, typename OutT>
__global__ void __launch_bounds__(256) k_grid_fwd(const __grid_constant__ GridP p, int64_t total, const float* __restrict__ x,
                                                  const RowT* __restrict__ table, OutT* __restrict__ y) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int64_t s = t / p.L;
    const int l = (int)(t - s * p.L);
    const float px = __ldg(x + 3 * s), py = __ldg(x + 3 * s + 1), pz = __ldg(x + 3 * s + 2);
    const Corner c = make_corner(px, py, pz, p.scale[l]);
    const uint32_t mask = (1u << p.log2T) - 1u;
    const RowT* slab = table + ((size_t)l << p.log2T);
    // issue all gathers before any use
    float2 f[8];
    gather_level(slab, c, mask, f);
    const float mx = __fsub_rn(1.f, c.ox), my = __fsub_rn(1.f, c.oy), mz = __fsub_rn(1.f, c.oz);
    float out[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#define FJ(k) (j == 0 ? f[k].x : f[k].y)
        const float f03 = lerp_ref(FJ(0), FJ(3), c.ox, mx);
        const float f12 = lerp_ref(FJ(1), FJ(2), c.ox, mx);
        const float f56 = lerp_ref(FJ(5), FJ(6), c.ox, mx);
        const float f47 = lerp_ref(FJ(4), FJ(7), c.ox, mx);
        const float f0312 = lerp_ref(f03, f12, c.oy, my);
        const float f4756 = lerp_ref(f47, f56, c.oy, my);
        out[j] = lerp_ref(f0312, f4756, c.oz, mz);
#undef FJ
    }
    store_feat(y, t, out[0], out[1]);
}

// Forward straight into the tensor-core MLP's operand layout (TMH: tiles of 128 rows, [chunk = col/8][row][8 halfs], see
// mlp_tc.cu): one thread per (row, chunk) computes the chunk's 4 levels (32 independent gathers in flight) and writes ONE
// 16-byte vector; consecutive lanes are consecutive rows of the same chunk, so a warp stores 512 contiguous bytes and the
// coarse-level gathers of neighbouring samples coalesce in L1.  Rows >= n and levels >= L are written as zeros (the MLP's
// zero padding), so the buffer needs no separate initialisation.
template <typename RowT>
__global__ void __launch_bounds__(256, 3) k_grid_fwd_tmh(const __grid_constant__ GridP p, int64_t n, int nch, const float* __restrict__ x,
                                                      const RowT* __restrict__ table, uint4* __restrict__ y) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= ((n + 127) >> 7 << 7)) return;  // beyond the last (padded) tile
    const int c = blockIdx.y;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (row < n) {
        const float px = __ldg(x + 3 * row), py = __ldg(x + 3 * row + 1), pz = __ldg(x + 3 * row + 2);
        const uint32_t mask = (1u << p.log2T) - 1u;
        float2 f[4][8];
        Corner cs[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int l = c * 4 + q;
            if (l < p.L) {
                cs[q] = make_corner(px, py, pz, p.scale[l]);
                gather_level(table + ((size_t)l << p.log2T), cs[q], mask, f[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int l = c * 4 + q;
            if (l < p.L) {
                const float2 o = trilerp_ref(f[q], cs[q]);
                v[2 * q] = o.x;
                v[2 * q + 1] = o.y;
            }
        }
    }
    const int64_t tile = row >> 7;
    const int r = (int)(row & 127);
    __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]), h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2);
    u.w = *reinterpret_cast<uint32_t*>(&h3);
    y[(tile * nch + c) * 128 + r] = u;
}

// The same forward, additionally saving d(feature)/d(x) so that dL/dx later costs no second gather pass (what tcnn's forward does
// with `dy_dx` when input gradients are prepared, TCNN/.../encodings/grid.h:160-211; nerfstudio needs dL/dx every step for the
// density-gradient normals, NS/fields/base_field.py:80-101).  Per (row, chunk) three more 16-byte vectors: the x / y / z derivatives
// of the chunk's 8 features in UNSCALED form (derivative w.r.t. the level's scaled coordinate; the consumer multiplies by scale_l,
// which keeps the fp16 range independent of the level resolution).  Layout: [tile][chunk][axis][row 0..127][8 halfs].
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
template <typename RowT, bool PAIR>
__global__ void __launch_bounds__(256, 2) k_grid_fwd_tmh_jac(const __grid_constant__ GridP p, int64_t n, int nch, const float* __restrict__ x,
                                                          const RowT* __restrict__ table, uint4* __restrict__ y, uint4* __restrict__ jac) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= ((n + 127) >> 7 << 7)) return;  // beyond the last (padded) tile
    const int c = blockIdx.y;
    uint32_t v[4] = {0u, 0u, 0u, 0u}, jx[4] = {0u, 0u, 0u, 0u}, jy[4] = {0u, 0u, 0u, 0u}, jz[4] = {0u, 0u, 0u, 0u};
    if (row < n) {
        const float px = __ldg(x + 3 * row), py = __ldg(x + 3 * row + 1), pz = __ldg(x + 3 * row + 2);
        const uint32_t mask = (1u << p.log2T) - 1u;
        float2 f[4][8];
        Corner cs[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int l = c * 4 + q;
            if (l < p.L) {
                cs[q] = make_corner(px, py, pz, p.scale[l]);
                gather_level<RowT, PAIR>(table + ((size_t)l << p.log2T), cs[q], mask, f[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int l = c * 4 + q;
            if (l < p.L) {
                const float2 o = trilerp_ref(f[q], cs[q]);  // the feature itself: the reference's rounding sequence, as k_grid_fwd_tmh
                v[q] = pack_h2(o.x, o.y);
                const float ox = cs[q].ox, oy = cs[q].oy, oz = cs[q].oz;
                const float mx = 1.f - ox, my = 1.f - oy, mz = 1.f - oz;
                float d[2][3];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
#define FJ(k) (j == 0 ? f[q][k].x : f[q][k].y)
                    const float f03 = FJ(0) * ox + FJ(3) * mx, f12 = FJ(1) * ox + FJ(2) * mx;
                    const float f56 = FJ(5) * ox + FJ(6) * mx, f47 = FJ(4) * ox + FJ(7) * mx;
                    d[j][2] = (f03 * oy + f12 * my) - (f47 * oy + f56 * my);
                    d[j][1] = oz * (f03 - f12) + mz * (f47 - f56);
                    d[j][0] = oz * (oy * (FJ(0) - FJ(3)) + my * (FJ(1) - FJ(2))) + mz * (oy * (FJ(4) - FJ(7)) + my * (FJ(5) - FJ(6)));
#undef FJ
                }
                jx[q] = pack_h2(d[0][0], d[1][0]);
                jy[q] = pack_h2(d[0][1], d[1][1]);
                jz[q] = pack_h2(d[0][2], d[1][2]);
            }
        }
    }
    const int64_t tile = row >> 7;
    const int r = (int)(row & 127);
    y[(tile * nch + c) * 128 + r] = make_uint4(v[0], v[1], v[2], v[3]);
    uint4* jb = jac + ((tile * nch + c) * 3) * 128 + r;
    jb[0] = make_uint4(jx[0], jx[1], jx[2], jx[3]);
    jb[128] = make_uint4(jy[0], jy[1], jy[2], jy[3]);
    jb[256] = make_uint4(jz[0], jz[1], jz[2], jz[3]);
}

// dL/dx from the saved derivatives: dx[row][axis] = sum_l scale_l * (J[l,0,axis] * dy[2l] + J[l,1,axis] * dy[2l+1]); dy is the fp32
// tile-major buffer of the tensor-core MLP backward ([tile][2L][128]).  One thread per row: every access of a warp is contiguous.
// normalize_scale != 0: the result is written as normalize_scale * v / max(|v|, eps) (the normals epilogue, base_field.py:97-99).
__global__ void __launch_bounds__(128) k_grid_jac_dx(const __grid_constant__ GridP p, int64_t n, int nch, const uint4* __restrict__ jac,
                                                     const float* __restrict__ dy, float normalize_scale, float eps, float* __restrict__ dx) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const int64_t tile = row >> 7;
    const int r = (int)(row & 127);
    float acc[3] = {0.f, 0.f, 0.f};
    const int nreal = (p.L + 3) >> 2;
    for (int c = 0; c < nreal; ++c) {
        const uint4* jb = jac + ((tile * nch + c) * 3) * 128 + r;
        const uint4 J[3] = {__ldg(jb), __ldg(jb + 128), __ldg(jb + 256)};
        const float* db = dy + (((tile * (2 * p.L)) + 8 * c) << 7) + r;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int l = c * 4 + q;
            if (l < p.L) {
                const float s = p.scale[l];
                const float g0 = __ldg(db + ((2 * q) << 7)) * s, g1 = __ldg(db + ((2 * q + 1) << 7)) * s;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const uint32_t w = q == 0 ? J[a].x : q == 1 ? J[a].y : q == 2 ? J[a].z : J[a].w;
                    const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&w));
                    acc[a] = fmaf(d.x, g0, fmaf(d.y, g1, acc[a]));
                }
            }
        }
    }
    if (normalize_scale != 0.f) {
        const float nrm = fmaxf(sqrtf(acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2]), eps);
#pragma unroll
        for (int a = 0; a < 3; ++a) acc[a] = normalize_scale * (acc[a] / nrm);
    }
    dx[3 * row] = acc[0];
    dx[3 * row + 1] = acc[1];
    dx[3 * row + 2] = acc[2];
}

// Backward scatter: one thread walks GB = 4 consecutive samples (neighbours along a ray) through a group of 4 levels, merging
// equal-cell runs in registers and pairing the x-floor / x-ceil rows into 16-byte reductions (CellRun, grid_common.cuh).
// blockIdx.y = level group, so all 16 levels of a sample quad are in flight on different CTAs.
#define GB 4
template <typename OutT>
__global__ void __launch_bounds__(128) k_grid_bwd(const __grid_constant__ GridP p, int64_t n, const float* __restrict__ x,
                                                  const OutT* __restrict__ dy, float* __restrict__ dtable, int tmf) {
    const int64_t t0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * GB;
    if (t0 >= n) return;
    float q[GB][3];
    if (t0 + GB <= n && (reinterpret_cast<size_t>(x) & 15) == 0) {  // 12 contiguous floats, 48-byte aligned
        const float4* xv = reinterpret_cast<const float4*>(x + 3 * t0);
        const float4 a = __ldg(xv), b = __ldg(xv + 1), c = __ldg(xv + 2);
        q[0][0] = a.x, q[0][1] = a.y, q[0][2] = a.z, q[1][0] = a.w, q[1][1] = b.x, q[1][2] = b.y;
        q[2][0] = b.z, q[2][1] = b.w, q[2][2] = c.x, q[3][0] = c.y, q[3][1] = c.z, q[3][2] = c.w;
    } else {
#pragma unroll
        for (int g = 0; g < GB; ++g) {
            const int64_t t = min(t0 + g, n - 1);
            q[g][0] = __ldg(x + 3 * t), q[g][1] = __ldg(x + 3 * t + 1), q[g][2] = __ldg(x + 3 * t + 2);
        }
    }
    const uint32_t mask = (1u << p.log2T) - 1u;
    const bool tmf_vec = tmf && sizeof(OutT) == 4 && (reinterpret_cast<size_t>(dy) & 15) == 0;
    const int l_end = min(p.L, (int)(blockIdx.y + 1) * 4);
    for (int l = blockIdx.y * 4; l < l_end; ++l) {
        float* slab = dtable + (((size_t)l << p.log2T) << 1);
        const float scale = p.scale[l];
        CellRun run;
        run.reset();
        float g0[GB], g1[GB];
        if (tmf_vec) {
            // fp32 tile-major dy: the quad's four samples are 16 contiguous bytes per feature column (t0 is a multiple of 4, tiles are 128 rows)
            const float* b = reinterpret_cast<const float*>(dy) + (((t0 >> 7) * (2 * p.L) + 2 * l) << 7) + (t0 & 127);
            const float4 a = __ldg(reinterpret_cast<const float4*>(b)), c = __ldg(reinterpret_cast<const float4*>(b + 128));
            g0[0] = a.x, g0[1] = a.y, g0[2] = a.z, g0[3] = a.w;
            g1[0] = c.x, g1[1] = c.y, g1[2] = c.z, g1[3] = c.w;
        } else {
#pragma unroll
            for (int g = 0; g < GB; ++g) {
                const float2 gr = load_dy(dy, min(t0 + g, n - 1), l, p.L, tmf);
                g0[g] = gr.x, g1[g] = gr.y;
            }
        }
#pragma unroll
        for (int g = 0; g < GB; ++g) {
            if (t0 + g >= n) break;
            if (g0[g] == 0.f && g1[g] == 0.f) continue;  // adding zeros changes nothing (masked / padded samples)
            run.add(slab, q[g][0], q[g][1], q[g][2], scale, mask, g0[g], g1[g]);
        }
        run.flush(slab);
    }
}

// Long-run variant for the tile-major fp32 dy of the tensor-core path: one thread walks G consecutive samples (G = 8 or 16: a third of
// a 48-sample ray) through ONE level (blockIdx.y), so equal-cell runs of the coarse levels — a ray's samples cluster around a surface —
// are merged over 2-4 times more samples before a reduction is issued.  Same thread count as k_grid_bwd's (quad, 4-level group).
template <int G>
__global__ void __launch_bounds__(128) k_grid_bwd_run(const __grid_constant__ GridP p, int64_t n, const float* __restrict__ x,
                                                      const float* __restrict__ dy, float* __restrict__ dtable) {
    const int64_t t0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * G;
    if (t0 >= n) return;
    const int l = blockIdx.y;
    float* slab = dtable + (((size_t)l << p.log2T) << 1);
    const float scale = p.scale[l];
    const uint32_t mask = (1u << p.log2T) - 1u;
    // the G samples sit in one 128-row tile (G divides 128): G contiguous floats per feature column
    const float* b = dy + (((t0 >> 7) * (2 * p.L) + 2 * l) << 7) + (t0 & 127);
    float g0[G], g1[G];
#pragma unroll
    for (int k = 0; k < G; k += 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(b + k)), c = __ldg(reinterpret_cast<const float4*>(b + 128 + k));
        g0[k] = a.x, g0[k + 1] = a.y, g0[k + 2] = a.z, g0[k + 3] = a.w;
        g1[k] = c.x, g1[k + 1] = c.y, g1[k + 2] = c.z, g1[k + 3] = c.w;
    }
    CellRun run;
    run.reset();
    const bool full = t0 + G <= n;
#pragma unroll
    for (int k = 0; k < G; k += 4) {
        float q[4][3];
        if (full) {  // x + 3*t0 is 16-byte aligned: t0 is a multiple of 4 and the caller checked the base pointer
            const float4* xv = reinterpret_cast<const float4*>(x + 3 * (t0 + k));
            const float4 a = __ldg(xv), bb = __ldg(xv + 1), c = __ldg(xv + 2);
            q[0][0] = a.x, q[0][1] = a.y, q[0][2] = a.z, q[1][0] = a.w, q[1][1] = bb.x, q[1][2] = bb.y;
            q[2][0] = bb.z, q[2][1] = bb.w, q[2][2] = c.x, q[3][0] = c.y, q[3][1] = c.z, q[3][2] = c.w;
        } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int64_t t = min(t0 + k + g, n - 1);
                q[g][0] = __ldg(x + 3 * t), q[g][1] = __ldg(x + 3 * t + 1), q[g][2] = __ldg(x + 3 * t + 2);
            }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (t0 + k + g >= n) break;
            if (g0[k + g] == 0.f && g1[k + g] == 0.f) continue;
            run.add(slab, q[g][0], q[g][1], q[g][2], scale, mask, g0[k + g], g1[k + g]);
        }
    }
    run.flush(slab);
}

// Rolled variant of k_grid_bwd_run: the same walk (G consecutive samples of a ray, one level, run merging in registers), but the G / 4 sample
// quads are a real loop with the next quad's positions and gradients requested before the current one is processed.  The fully unrolled
// kernel keeps all 2 G gradient values live and expands CellRun's flush G + 1 times (3 584 instructions, 92 registers: 5 CTAs per SM,
// instruction-cache misses on the 5-level proposal launches); this one is a quarter of the code and fits 6 CTAs per SM.
template <int G>
__global__ void __launch_bounds__(128, 6) k_grid_bwd_run_rolled(const __grid_constant__ GridP p, int64_t n, const float* __restrict__ x,
                                                                const float* __restrict__ dy, float* __restrict__ dtable) {
    const int64_t t0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * G;
    if (t0 >= n) return;
    const int l = blockIdx.y;
    float* slab = dtable + (((size_t)l << p.log2T) << 1);
    const float scale = p.scale[l];
    const uint32_t mask = (1u << p.log2T) - 1u;
    const float* b = dy + (((t0 >> 7) * (2 * p.L) + 2 * l) << 7) + (t0 & 127);
    const bool full = t0 + G <= n;
    CellRun run;
    run.reset();
    if (!full) {  // ragged tail of the batch: scalar walk
        for (int k = 0; k < G && t0 + k < n; ++k) {
            const float g0 = __ldg(b + k), g1 = __ldg(b + 128 + k);
            if (g0 == 0.f && g1 == 0.f) continue;
            const int64_t t = t0 + k;
            run.add(slab, __ldg(x + 3 * t), __ldg(x + 3 * t + 1), __ldg(x + 3 * t + 2), scale, mask, g0, g1);
        }
        run.flush(slab);
        return;
    }
    const float4* xv = reinterpret_cast<const float4*>(x + 3 * t0);  // 16-byte aligned: t0 is a multiple of 4 and the caller checked the base
    float4 na = __ldg(reinterpret_cast<const float4*>(b)), nc = __ldg(reinterpret_cast<const float4*>(b + 128));
    float4 nx0 = __ldg(xv), nx1 = __ldg(xv + 1), nx2 = __ldg(xv + 2);
#pragma unroll 1
    for (int k = 0; k < G; k += 4) {
        const float4 a = na, c = nc, x0 = nx0, x1 = nx1, x2 = nx2;
        if (k + 4 < G) {
            na = __ldg(reinterpret_cast<const float4*>(b + k + 4)), nc = __ldg(reinterpret_cast<const float4*>(b + 128 + k + 4));
            nx0 = __ldg(xv + 3 * (k / 4 + 1)), nx1 = __ldg(xv + 3 * (k / 4 + 1) + 1), nx2 = __ldg(xv + 3 * (k / 4 + 1) + 2);
        }
        if (!(a.x == 0.f && c.x == 0.f)) run.add(slab, x0.x, x0.y, x0.z, scale, mask, a.x, c.x);
        if (!(a.y == 0.f && c.y == 0.f)) run.add(slab, x0.w, x1.x, x1.y, scale, mask, a.y, c.y);
        if (!(a.z == 0.f && c.z == 0.f)) run.add(slab, x1.z, x1.w, x2.x, scale, mask, a.z, c.z);
        if (!(a.w == 0.f && c.w == 0.f)) run.add(slab, x2.y, x2.z, x2.w, scale, mask, a.w, c.w);
    }
    run.flush(slab);
}

// dL/dx: one thread per (sample, level) computes its level's contribution, then the L lanes of a sample are
// summed with warp shuffles when L divides 32 (main grid, L=16), else with atomics (L=5 proposals).
template <typename RowT, typename OutT>
__global__ void __launch_bounds__(256) k_grid_bwd_input(const __grid_constant__ GridP p, int64_t total, const float* __restrict__ x,
                                                        const RowT* __restrict__ table, const OutT* __restrict__ dy, float* __restrict__ dx,
                                                        int shuffle_reduce, int tmf) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < total;
    const int64_t tt = live ? t : total - 1;
    const int64_t s = tt / p.L;
    const int l = (int)(tt - s * p.L);
    const float scale = p.scale[l];
    const float px = __ldg(x + 3 * s), py = __ldg(x + 3 * s + 1), pz = __ldg(x + 3 * s + 2);
    const Corner c = make_corner(px, py, pz, scale);
    const uint32_t mask = (1u << p.log2T) - 1u;
    const RowT* slab = table + ((size_t)l << p.log2T);
    float2 f[8];
    gather_level(slab, c, mask, f);
    float2 g = load_dy(dy, s, l, p.L, tmf);
    if (!live) g = make_float2(0.f, 0.f);
    const float mx = 1.f - c.ox, my = 1.f - c.oy, mz = 1.f - c.oz;
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#define FJ(k) (j == 0 ? f[k].x : f[k].y)
        const float gj = j == 0 ? g.x : g.y;
        const float f03 = FJ(0) * c.ox + FJ(3) * mx, f12 = FJ(1) * c.ox + FJ(2) * mx;
        const float f56 = FJ(5) * c.ox + FJ(6) * mx, f47 = FJ(4) * c.ox + FJ(7) * mx;
        const float f0312 = f03 * c.oy + f12 * my, f4756 = f47 * c.oy + f56 * my;
        // d/d oz, d/d oy, d/d ox of the nested lerp
        gz += gj * (f0312 - f4756);
        gy += gj * (c.oz * (f03 - f12) + mz * (f47 - f56));
        gx += gj * (c.oz * (c.oy * (FJ(0) - FJ(3)) + my * (FJ(1) - FJ(2))) + mz * (c.oy * (FJ(4) - FJ(7)) + my * (FJ(5) - FJ(6))));
#undef FJ
    }
    gx *= scale;
    gy *= scale;
    gz *= scale;
    if (shuffle_reduce) {
        // L is a power of two dividing 32: lanes [k*L, (k+1)*L) hold one sample
        for (int o = p.L >> 1; o > 0; o >>= 1) {
            gx += __shfl_xor_sync(0xffffffffu, gx, o);
            gy += __shfl_xor_sync(0xffffffffu, gy, o);
            gz += __shfl_xor_sync(0xffffffffu, gz, o);
        }
        if (live && l == 0) {
            dx[3 * s] = gx;
            dx[3 * s + 1] = gy;
            dx[3 * s + 2] = gz;
        }
    } else if (live) {
        atomicAdd(dx + 3 * s, gx);
        atomicAdd(dx + 3 * s + 1, gy);
        atomicAdd(dx + 3 * s + 2, gz);
    }
}

// dL/dx, one thread per SAMPLE walking all LL levels (the proposal grids' L = 5 does not divide a warp: the (sample, level) kernel above
// then needs a zero fill and three atomics per thread): 8 LL gathers in flight per thread, one 12-byte store.
template <int LL>
__global__ void __launch_bounds__(256) k_grid_bwd_input_sample(const __grid_constant__ GridP p, int64_t n, const float* __restrict__ x,
                                                               const float2* __restrict__ table, const float* __restrict__ dy, float* __restrict__ dx, int tmf) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float px = __ldg(x + 3 * s), py = __ldg(x + 3 * s + 1), pz = __ldg(x + 3 * s + 2);
    const uint32_t mask = (1u << p.log2T) - 1u;
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
    for (int l = 0; l < LL; ++l) {
        const float scale = p.scale[l];
        const Corner c = make_corner(px, py, pz, scale);
        float2 f[8];
        gather_level(table + ((size_t)l << p.log2T), c, mask, f);
        const float2 g = load_dy(dy, s, l, LL, tmf);
        const float mx = 1.f - c.ox, my = 1.f - c.oy, mz = 1.f - c.oz;
        float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#define FJ(k) (j == 0 ? f[k].x : f[k].y)
            const float gj = j == 0 ? g.x : g.y;
            const float f03 = FJ(0) * c.ox + FJ(3) * mx, f12 = FJ(1) * c.ox + FJ(2) * mx;
            const float f56 = FJ(5) * c.ox + FJ(6) * mx, f47 = FJ(4) * c.ox + FJ(7) * mx;
            const float f0312 = f03 * c.oy + f12 * my, f4756 = f47 * c.oy + f56 * my;
            az += gj * (f0312 - f4756);
            ay += gj * (c.oz * (f03 - f12) + mz * (f47 - f56));
            ax += gj * (c.oz * (c.oy * (FJ(0) - FJ(3)) + my * (FJ(1) - FJ(2))) + mz * (c.oy * (FJ(4) - FJ(7)) + my * (FJ(5) - FJ(6))));
#undef FJ
        }
        gx += ax * scale, gy += ay * scale, gz += az * scale;
    }
    dx[3 * s] = gx, dx[3 * s + 1] = gy, dx[3 * s + 2] = gz;
}

__global__ void __launch_bounds__(256) k_grid_indices(const __grid_constant__ GridP p, int64_t total, const float* __restrict__ x, int64_t* __restrict__ idx) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int64_t s = t / p.L;
    const int l = (int)(t - s * p.L);
    const Corner c = make_corner(x[3 * s], x[3 * s + 1], x[3 * s + 2], p.scale[l]);
    const uint32_t mask = (1u << p.log2T) - 1u;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        idx[t * 8 + k] = (int64_t)corner_index(c, SEL_X(k), SEL_Y(k), SEL_Z(k), mask) + ((int64_t)l << p.log2T);
}

static int make_params(const nvo_grid_desc* d, GridP* p) {
    NVO_CHECK(d != nullptr, "grid: null descriptor");
    NVO_CHECK(d->n_levels >= 1 && d->n_levels <= NVO_MAX_LEVELS, "grid: n_levels=%d out of range [1,%d]", d->n_levels, NVO_MAX_LEVELS);
    NVO_CHECK(d->log2_T >= 1 && d->log2_T <= 30, "grid: log2_T=%d out of range [1,30]", d->log2_T);
    NVO_CHECK(d->table_dtype == NVO_F32 || d->table_dtype == NVO_F16, "grid: bad table_dtype %d", d->table_dtype);
    NVO_CHECK(d->out_dtype >= NVO_F32 && d->out_dtype <= NVO_F32_TMF, "grid: bad out_dtype %d", d->out_dtype);
    p->L = d->n_levels;
    p->log2T = d->log2_T;
    for (int i = 0; i < NVO_MAX_LEVELS; ++i) p->scale[i] = i < d->n_levels ? d->scalings[i] : 0.f;
    return 0;
}

extern "C" int nvo_grid_forward(const nvo_grid_desc* d, void* stream, int64_t n, const float* x, const void* table, void* y) {
    GridP p;
    if (int e = make_params(d, &p)) return e;
    NVO_CHECK(n >= 0, "grid_forward: negative batch");
    if (n == 0) return 0;
    NVO_CHECK(x && table && y, "grid_forward: null pointer");
    const int64_t total = n * p.L;
    cudaStream_t st = (cudaStream_t)stream;
    if (d->out_dtype == NVO_F16_TMH) {
        const int nch = ((2 * p.L + 15) & ~15) >> 3;  // feature columns padded to the MMA K granularity (16)
        const dim3 grid((unsigned int)(((n + 127) / 128 * 128 + 255) / 256), (unsigned int)nch);
        if (d->table_dtype == NVO_F32)
            k_grid_fwd_tmh<float2><<<grid, 256, 0, st>>>(p, n, nch, x, (const float2*)table, (uint4*)y);
        else
            k_grid_fwd_tmh<__half2><<<grid, 256, 0, st>>>(p, n, nch, x, (const __half2*)table, (uint4*)y);
        NVO_CUDA_LAUNCH_CHECK("grid_forward(tmh)");
        return 0;
    }
    const unsigned int g = nvo_blocks(total, 256);
    if (d->table_dtype == NVO_F32 && d->out_dtype == NVO_F32)
        k_grid_fwd<float2, float><<<g, 256, 0, st>>>(p, total, x, (const float2*)table, (float*)y);
    else if (d->table_dtype == NVO_F32)
        k_grid_fwd<float2, __half><<<g, 256, 0, st>>>(p, total, x, (const float2*)table, (__half*)y);
    else if (d->out_dtype == NVO_F32)
        k_grid_fwd<__half2, float><<<g, 256, 0, st>>>(p, total, x, (const __half2*)table, (float*)y);
    else
        k_grid_fwd<__half2, __half><<<g, 256, 0, st>>>(p, total, x, (const __half2*)table, (__half*)y);
    NVO_CUDA_LAUNCH_CHECK("grid_forward");
    return 0;
}

extern "C" int nvo_grid_forward_jac(const nvo_grid_desc* d, void* stream, int64_t n, const float* x, const void* table, void* y, void* jac) {
    GridP p;
    if (int e = make_params(d, &p)) return e;
    NVO_CHECK(n >= 0, "grid_forward_jac: negative batch");
    NVO_CHECK(d->out_dtype == NVO_F16_TMH, "grid_forward_jac: output layout must be NVO_F16_TMH");
    if (n == 0) return 0;
    NVO_CHECK(x && table && y && jac, "grid_forward_jac: null pointer");
    const int nch = ((2 * p.L + 15) & ~15) >> 3;
    const dim3 grid((unsigned int)(((n + 127) / 128 * 128 + 255) / 256), (unsigned int)nch);
    cudaStream_t st = (cudaStream_t)stream;
    static const int pair = nvo_env_int("NVO_GRID_FWD_PAIR", 1);  // 0: eight plain gathers per level instead of paired 16-byte loads behind a branch
    if (d->table_dtype == NVO_F32) {
        if (pair)
            k_grid_fwd_tmh_jac<float2, true><<<grid, 256, 0, st>>>(p, n, nch, x, (const float2*)table, (uint4*)y, (uint4*)jac);
        else
            k_grid_fwd_tmh_jac<float2, false><<<grid, 256, 0, st>>>(p, n, nch, x, (const float2*)table, (uint4*)y, (uint4*)jac);
    } else
        k_grid_fwd_tmh_jac<__half2, true><<<grid, 256, 0, st>>>(p, n, nch, x, (const __half2*)table, (uint4*)y, (uint4*)jac);
    NVO_CUDA_LAUNCH_CHECK("grid_forward_jac");
    return 0;
}

extern "C" int nvo_grid_jac_dx(const nvo_grid_desc* d, void* stream, int64_t n, const void* jac, const float* dy, float normalize_scale, float eps,
                               float* dx) {
    GridP p;
    if (int e = make_params(d, &p)) return e;
    NVO_CHECK(n >= 0, "grid_jac_dx: negative batch");
    if (n == 0) return 0;
    NVO_CHECK(jac && dy && dx, "grid_jac_dx: null pointer");
    const int nch = ((2 * p.L + 15) & ~15) >> 3;
    k_grid_jac_dx<<<nvo_blocks(n, 128), 128, 0, (cudaStream_t)stream>>>(p, n, nch, (const uint4*)jac, dy, normalize_scale, eps, dx);
    NVO_CUDA_LAUNCH_CHECK("grid_jac_dx");
    return 0;
}

extern "C" int nvo_grid_backward(const nvo_grid_desc* d, void* stream, int64_t n, const float* x, const void* dy, float* dtable) {
    GridP p;
    if (int e = make_params(d, &p)) return e;
    NVO_CHECK(n >= 0, "grid_backward: negative batch");
    if (n == 0) return 0;
    NVO_CHECK(x && dy && dtable, "grid_backward: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (d->out_dtype == NVO_F32_TMF && (reinterpret_cast<size_t>(dy) & 15) == 0 && (reinterpret_cast<size_t>(x) & 15) == 0) {
        // samples per thread and level: 16 (default; measured on a step's own sample positions: 146 -> 121 us), 8, or 0 = the quad kernel
        static const int run_g = nvo_env_int("NVO_GRID_BWD_RUN", 16);  // read once per process
        // rolled walk (k_grid_bwd_run_rolled): 0 = off, 1 = grids of up to 8 levels (the proposal networks'), 2 = every grid
        static const int rolled = nvo_env_int("NVO_GRID_BWD_ROLLED", 0);
        static const int rolled_g = nvo_env_int("NVO_GRID_BWD_ROLLED_G", 16);
        if (rolled == 2 || (rolled == 1 && p.L <= 8)) {
            const int g = rolled_g == 8 ? 8 : rolled_g == 32 ? 32 : 16;
            const dim3 gr(nvo_blocks((n + g - 1) / g, 128), (unsigned int)p.L);
            if (g == 8)
                k_grid_bwd_run_rolled<8><<<gr, 128, 0, st>>>(p, n, x, (const float*)dy, dtable);
            else if (g == 32)
                k_grid_bwd_run_rolled<32><<<gr, 128, 0, st>>>(p, n, x, (const float*)dy, dtable);
            else
                k_grid_bwd_run_rolled<16><<<gr, 128, 0, st>>>(p, n, x, (const float*)dy, dtable);
            NVO_CUDA_LAUNCH_CHECK("grid_backward(rolled)");
            return 0;
        }
        if (run_g == 8 || run_g == 16) {
            const dim3 gr(nvo_blocks((n + run_g - 1) / run_g, 128), (unsigned int)p.L);
            if (run_g == 8)
                k_grid_bwd_run<8><<<gr, 128, 0, st>>>(p, n, x, (const float*)dy, dtable);
            else
                k_grid_bwd_run<16><<<gr, 128, 0, st>>>(p, n, x, (const float*)dy, dtable);
            NVO_CUDA_LAUNCH_CHECK("grid_backward(run)");
            return 0;
        }
    }
    const dim3 g(nvo_blocks((n + GB - 1) / GB, 128), (unsigned int)((p.L + 3) / 4));
    if (d->out_dtype == NVO_F32 || d->out_dtype == NVO_F32_TMF)
        k_grid_bwd<float><<<g, 128, 0, st>>>(p, n, x, (const float*)dy, dtable, d->out_dtype == NVO_F32_TMF);
    else
        k_grid_bwd<__half><<<g, 128, 0, st>>>(p, n, x, (const __half*)dy, dtable, 0);
    NVO_CUDA_LAUNCH_CHECK("grid_backward");
    return 0;
}

extern "C" int nvo_grid_backward_input(const nvo_grid_desc* d, void* stream, int64_t n, const float* x, const void* table, const void* dy, float* dx) {
    GridP p;
    if (int e = make_params(d, &p)) return e;
    NVO_CHECK(n >= 0, "grid_backward_input: negative batch");
    if (n == 0) return 0;
    NVO_CHECK(x && table && dy && dx, "grid_backward_input: null pointer");
    const int64_t total = n * p.L;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned int g = nvo_blocks(total, 256);
    const int shuffle = (p.L <= 32 && (p.L & (p.L - 1)) == 0) ? 1 : 0;
    if (p.L == 5 && d->table_dtype == NVO_F32 && (d->out_dtype == NVO_F32 || d->out_dtype == NVO_F32_TMF)) {
        k_grid_bwd_input_sample<5><<<nvo_blocks(n, 256), 256, 0, st>>>(p, n, x, (const float2*)table, (const float*)dy, dx, d->out_dtype == NVO_F32_TMF);
        NVO_CUDA_LAUNCH_CHECK("grid_backward_input(sample)");
        return 0;
    }
    if (!shuffle) {
        cudaError_t e = cudaMemsetAsync(dx, 0, sizeof(float) * 3 * n, st);
        NVO_CHECK(e == cudaSuccess, "grid_backward_input: memset failed: %s", cudaGetErrorString(e));
    }
    const int tmf = d->out_dtype == NVO_F32_TMF;
    const bool f32 = d->out_dtype == NVO_F32 || tmf;
    if (d->table_dtype == NVO_F32 && f32)
        k_grid_bwd_input<float2, float><<<g, 256, 0, st>>>(p, total, x, (const float2*)table, (const float*)dy, dx, shuffle, tmf);
    else if (d->table_dtype == NVO_F32)
        k_grid_bwd_input<float2, __half><<<g, 256, 0, st>>>(p, total, x, (const float2*)table, (const __half*)dy, dx, shuffle, 0);
    else if (f32)
        k_grid_bwd_input<__half2, float><<<g, 256, 0, st>>>(p, total, x, (const __half2*)table, (const float*)dy, dx, shuffle, tmf);
    else
        k_grid_bwd_input<__half2, __half><<<g, 256, 0, st>>>(p, total, x, (const __half2*)table, (const __half*)dy, dx, shuffle, 0);
    NVO_CUDA_LAUNCH_CHECK("grid_backward_input");
    return 0;
}

extern "C" int nvo_grid_indices(const nvo_grid_desc* d, void* stream, int64_t n, const float* x, int64_t* idx) {
    GridP p;
    if (int e = make_params(d, &p)) return e;
    NVO_CHECK(n >= 0, "grid_indices: negative batch");
    if (n == 0) return 0;
    NVO_CHECK(x && idx, "grid_indices: null pointer");
    const int64_t total = n * p.L;
    k_grid_indices<<<nvo_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(p, total, x, idx);
    NVO_CUDA_LAUNCH_CHECK("grid_indices");
    return 0;
}
