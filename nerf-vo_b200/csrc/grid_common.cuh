// Device-side building blocks shared by the hash-grid kernels (grid.cu) and the fused proposal-field kernels (prop.cu).
// Semantics = the reference's torch path (NS/field_components/encodings.py:405-465): scaled = x * scale_l (fp32),
// corners floor/ceil, hash (x*1 ^ y*2654435761 ^ z*805459861) mod 2^k + l*2^k, interpolation order x, y, z with
// separately rounded mul/add (no FMA contraction) so that the fp32 path reproduces the reference bit for bit.
#pragma once
#include "nvo_common.cuh"

struct GridP {
    int L;
    int log2T;
    float scale[NVO_MAX_LEVELS];
};

#define PRIME_Y 2654435761u
#define PRIME_Z 805459861u

struct Corner {
    float ox, oy, oz;              // fractional offsets
    uint32_t hx[2], hy[2], hz[2];  // per-axis hash terms for floor (0) / ceil (1)
};

__device__ __forceinline__ Corner make_corner(float px, float py, float pz, float scale) {
    Corner c;
    const float sx = __fmul_rn(px, scale), sy = __fmul_rn(py, scale), sz = __fmul_rn(pz, scale);
    const float fx = floorf(sx), fy = floorf(sy), fz = floorf(sz);
    c.ox = __fsub_rn(sx, fx);
    c.oy = __fsub_rn(sy, fy);
    c.oz = __fsub_rn(sz, fz);
    const int ifx = (int)fx, ify = (int)fy, ifz = (int)fz;
    const int icx = (int)ceilf(sx), icy = (int)ceilf(sy), icz = (int)ceilf(sz);
    c.hx[0] = (uint32_t)ifx;
    c.hx[1] = (uint32_t)icx;
    c.hy[0] = (uint32_t)ify * PRIME_Y;
    c.hy[1] = (uint32_t)icy * PRIME_Y;
    c.hz[0] = (uint32_t)ifz * PRIME_Z;
    c.hz[1] = (uint32_t)icz * PRIME_Z;
    return c;
}

// reference corner k -> (x,y,z) picks ceil(1)/floor(0): encodings.py:435-442
//   k: 0=(c,c,c) 1=(c,f,c) 2=(f,f,c) 3=(f,c,c) 4=(c,c,f) 5=(c,f,f) 6=(f,f,f) 7=(f,c,f); bit k of each mask
#define SEL_X(k) ((0x33 >> (k)) & 1)
#define SEL_Y(k) ((0x99 >> (k)) & 1)
#define SEL_Z(k) ((0x0F >> (k)) & 1)

__device__ __forceinline__ uint32_t corner_index(const Corner& c, int sx, int sy, int sz, uint32_t mask) {
    return (c.hx[sx] ^ c.hy[sy] ^ c.hz[sz]) & mask;
}

__device__ __forceinline__ float2 load_row(const float2* t, size_t i) { return __ldg(t + i); }
__device__ __forceinline__ float2 load_row(const __half2* t, size_t i) { return __half22float2(__ldg(t + i)); }

// a*w + b*(1-w) with the reference's rounding sequence (mul, mul, add)
__device__ __forceinline__ float lerp_ref(float a, float b, float w, float omw) { return __fadd_rn(__fmul_rn(a, w), __fmul_rn(b, omw)); }

// trilinear interpolation of the 8 gathered rows in the reference's order (encodings.py:453-463)
__device__ __forceinline__ float2 trilerp_ref(const float2* f, const Corner& c) {
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
    return make_float2(out[0], out[1]);
}

// one level of the encoding for one sample: 8 gathers issued back to back, then the interpolation
template <typename RowT>
__device__ __forceinline__ float2 grid_level_forward(const RowT* __restrict__ slab, const Corner& c, uint32_t mask) {
    float2 f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = load_row(slab, corner_index(c, SEL_X(k), SEL_Y(k), SEL_Z(k), mask));
    return trilerp_ref(f, c);
}

// SceneContraction(order=inf) + (x+2)/4 + selector masking of one point (spatial_distortions.py:67-69,
// nerfacto_field.py:201-209); returns the selector (0/1), q = normalised position multiplied by it.
__device__ __forceinline__ float contract_point(const float* p, float* q) {
    const float mag = fmaxf(fabsf(p[0]), fmaxf(fabsf(p[1]), fabsf(p[2])));
    bool sel = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float c = p[a];
        if (!(mag < 1.f)) c = __fmul_rn(__fsub_rn(2.f, __fdiv_rn(1.f, mag)), __fdiv_rn(p[a], mag));
        q[a] = __fdiv_rn(__fadd_rn(c, 2.f), 4.f);
        sel = sel && (q[a] > 0.f) && (q[a] < 1.f);
    }
    const float m = sel ? 1.f : 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) q[a] = __fmul_rn(q[a], m);
    return m;
}

// ---------------------------------------------------------------------------------------------------------------------
// Warp-level pre-reduction of the gradient scatter.  The 32 lanes of a warp hold 32 CONSECUTIVE samples (along a ray),
// which at coarse levels fall into the same grid cell in long runs; one `red` per run instead of one per lane removes
// most of the same-address serialisation in the L2 atomic units.  Runs = maximal stretches of adjacent lanes whose
// target row index is equal; a segmented Hillis-Steele scan sums each run into its last lane, which issues the
// reduction.  The scan depth adapts to the longest run in the warp (0 steps when every lane hits a different row).
// All 32 lanes must call this (lanes without work pass valid = false and a = b = 0).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void seg_red_add_v2(float* __restrict__ slab, uint32_t idx, float a, float b, bool valid, int lane) {
    const uint32_t key = valid ? idx : (0x80000000u | (uint32_t)lane);  // invalid lanes never merge with anything
    const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != key);
    const unsigned cont = ~heads;  // bit i set: lane i continues the run of lane i-1
    if (cont != 0u) {
        const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
        const unsigned r1 = cont & (cont << 1), r2 = r1 & (r1 << 2), r3 = r2 & (r2 << 4), r4 = r3 & (r3 << 8);
        const int nsteps = 1 + (r1 != 0u) + (r2 != 0u) + (r3 != 0u) + (r4 != 0u);
        for (int k = 0, o = 1; k < nsteps; ++k, o <<= 1) {
            const float ta = __shfl_up_sync(0xffffffffu, a, o), tb = __shfl_up_sync(0xffffffffu, b, o);
            if (lane - o >= start) {
                a += ta;
                b += tb;
            }
        }
    }
    const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
    if (tail && valid && (a != 0.f || b != 0.f)) nvo_red_add_v2(slab + 2 * (size_t)idx, a, b);
}

// scatter of one (sample, level): dL/dy = (g0, g1) into the 8 corner rows with the trilinear weights
__device__ __forceinline__ void grid_level_scatter(float* __restrict__ slab, const Corner& c, uint32_t mask, float g0, float g1, bool valid, int lane) {
    const float wx[2] = {1.f - c.ox, c.ox}, wy[2] = {1.f - c.oy, c.oy}, wz[2] = {1.f - c.oz, c.oz};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int sx = SEL_X(k), sy = SEL_Y(k), sz = SEL_Z(k);
        const float w = wz[sz] * wy[sy] * wx[sx];
        seg_red_add_v2(slab, corner_index(c, sx, sy, sz, mask), g0 * w, g1 * w, valid, lane);
    }
}
