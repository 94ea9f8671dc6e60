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
    // ceil(s) == floor(s) + (s != floor(s)): the same integers as (int)ceilf(s) without a second round + convert
    const int icx = ifx + (sx != fx), icy = ify + (sy != fy), icz = ifz + (sz != fz);
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

// trilinear interpolation of the 8 gathered rows in the reference's order (encodings.py:453-463).  Both features of a row go through one packed
// instruction (sm_100 FMUL2 / FADD2: each half rounded like the scalar op, no contraction): 21 instead of 42 instructions per level, same bits.
// (The same packing in the scatter's CellRun measured neutral to slightly slower — 94 registers, extra moves — and was not kept.)
__device__ __forceinline__ float2 lerp_ref2(float2 a, float2 b, float2 w, float2 omw) { return __fadd2_rn(__fmul2_rn(a, w), __fmul2_rn(b, omw)); }
__device__ __forceinline__ float2 trilerp_ref(const float2* f, const Corner& c) {
    const float mx = __fsub_rn(1.f, c.ox), my = __fsub_rn(1.f, c.oy), mz = __fsub_rn(1.f, c.oz);
    const float2 ox2 = make_float2(c.ox, c.ox), mx2 = make_float2(mx, mx), oy2 = make_float2(c.oy, c.oy), my2 = make_float2(my, my);
    const float2 oz2 = make_float2(c.oz, c.oz), mz2 = make_float2(mz, mz);
    const float2 f03 = lerp_ref2(f[0], f[3], ox2, mx2);
    const float2 f12 = lerp_ref2(f[1], f[2], ox2, mx2);
    const float2 f56 = lerp_ref2(f[5], f[6], ox2, mx2);
    const float2 f47 = lerp_ref2(f[4], f[7], ox2, mx2);
    const float2 f0312 = lerp_ref2(f03, f12, oy2, my2);
    const float2 f4756 = lerp_ref2(f47, f56, oy2, my2);
    return lerp_ref2(f0312, f4756, oz2, mz2);
}

// Gather of the 8 corner rows of one (sample, level) into f[k] (reference corner numbering).  The hash is linear in x
// (prime 1), so for an even x-floor the x-floor / x-ceil rows of each (y,z) combination are r and r^1: ONE aligned double-width
// load fetches both.  Scattered gathers are bound by L1 tag lookups (one per distinct 32-byte sector per instruction), so this
// removes a quarter of them on average; the values are bit-identical to eight separate loads.
__device__ __forceinline__ void load_row_pair(const float2* t, uint32_t lo, float2& a, float2& b) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(t + lo));
    a = make_float2(v.x, v.y);
    b = make_float2(v.z, v.w);
}
__device__ __forceinline__ void load_row_pair(const __half2* t, uint32_t lo, float2& a, float2& b) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(t + lo));
    a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
    b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
}
// PAIR = false: eight plain loads, no divergent branch — for kernels that are issue-bound rather than bound by the gathers (the fused proposal
// forward: selectable with NVO_PROP_FWD_PAIR); the values are the same either way.
template <typename RowT, bool PAIR = true>
__device__ __forceinline__ void gather_level(const RowT* __restrict__ slab, const Corner& c, uint32_t mask, float2* f) {
    // reference corner k -> (x,y,z): 0=ccc 1=cfc 2=ffc 3=fcc 4=ccf 5=cff 6=fff 7=fcf.  For the (y,z) combination j = sy + 2*sz the
    // x-ceil member is corner kc[j] and the x-floor member corner kf[j]:  (f,f): 5,6   (c,f): 4,7   (f,c): 1,2   (c,c): 0,3
    const int kc[4] = {5, 4, 1, 0}, kf[4] = {6, 7, 2, 3};  // indexed by j = sy + 2*sz
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int sy = j & 1, sz = j >> 1;
        const uint32_t hyz = c.hy[sy] ^ c.hz[sz];
        const uint32_t idf = (c.hx[0] ^ hyz) & mask, idc = (c.hx[1] ^ hyz) & mask;
        if (PAIR && (idf ^ idc) == 1u) {
            float2 lo, hi;
            load_row_pair(slab, idf & idc, lo, hi);  // idf & idc == min(idf, idc): the even row of the pair
            f[kf[j]] = (idf & 1u) ? hi : lo;
            f[kc[j]] = (idf & 1u) ? lo : hi;
        } else {
            f[kf[j]] = load_row(slab, idf);
            f[kc[j]] = load_row(slab, idc);
        }
    }
}

// one level of the encoding for one sample: all gathers issued back to back, then the interpolation
template <typename RowT, bool PAIR = true>
__device__ __forceinline__ float2 grid_level_forward(const RowT* __restrict__ slab, const Corner& c, uint32_t mask) {
    float2 f[8];
    gather_level<RowT, PAIR>(slab, c, mask, f);
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
        q[a] = __fmul_rn(__fadd_rn(c, 2.f), 0.25f);
        sel = sel && (q[a] > 0.f) && (q[a] < 1.f);
    }
    const float m = sel ? 1.f : 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) q[a] = __fmul_rn(q[a], m);
    return m;
}

// ---------------------------------------------------------------------------------------------------------------------
// Sequential run-merging scatter (measured on B200, tools/atomics_probe.cu: L2 reductions retire at ~200 G requests/s whatever
// their width — f32, v2.f32, v4.f32 cost the same — lanes of one instruction hitting the SAME row are NOT merged by the
// hardware, while one red.v4.f32 updates two adjacent rows for the price of one).  So:
//   * a thread walks G consecutive samples of a ray and keeps, per (y,z) corner combination, the pending sum for the x-floor and
//     x-ceil rows in registers while the cell does not change (consecutive samples share coarse cells);
//   * on a change (and at the end) the pair is flushed: rows r and r^1 (x-floor even) go out as ONE 16-byte red.v4.f32,
//     otherwise as two red.v2.f32.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void nvo_red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct ScatterRun {
    uint32_t pf[4], pc[4];  // pending x-floor / x-ceil rows per (y,z) combination
    float af0[4], af1[4], ac0[4], ac1[4];
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int j = 0; j < 4; ++j) pf[j] = pc[j] = 0xffffffffu, af0[j] = af1[j] = ac0[j] = ac1[j] = 0.f;
    }
    __device__ __forceinline__ void flush(float* __restrict__ slab, int j) {
        const bool nf = af0[j] != 0.f || af1[j] != 0.f, nc = ac0[j] != 0.f || ac1[j] != 0.f;
        if (!nf && !nc) return;
        if ((pf[j] ^ pc[j]) == 1u) {  // adjacent rows of one aligned 16-byte pair
            if (pf[j] < pc[j])
                nvo_red_add_v4(slab + 2 * (size_t)pf[j], af0[j], af1[j], ac0[j], ac1[j]);
            else
                nvo_red_add_v4(slab + 2 * (size_t)pc[j], ac0[j], ac1[j], af0[j], af1[j]);
        } else {
            if (nf) nvo_red_add_v2(slab + 2 * (size_t)pf[j], af0[j], af1[j]);
            if (nc) nvo_red_add_v2(slab + 2 * (size_t)pc[j], ac0[j], ac1[j]);
        }
    }
    // add one sample's (g0, g1) at corner geometry c
    __device__ __forceinline__ void add(float* __restrict__ slab, const Corner& c, uint32_t mask, float g0, float g1) {
        const float wx0 = 1.f - c.ox, wx1 = c.ox;
        const float wy[2] = {1.f - c.oy, c.oy}, wz[2] = {1.f - c.oz, c.oz};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int sy = j & 1, sz = j >> 1;
            const uint32_t hyz = c.hy[sy] ^ c.hz[sz];
            const uint32_t idf = (c.hx[0] ^ hyz) & mask, idc = (c.hx[1] ^ hyz) & mask;
            if (idf != pf[j] || idc != pc[j]) {
                flush(slab, j);
                pf[j] = idf;
                pc[j] = idc;
                af0[j] = af1[j] = ac0[j] = ac1[j] = 0.f;
            }
            const float wyz = wz[sz] * wy[sy];
            const float wf = wyz * wx0, wc = wyz * wx1;
            af0[j] = fmaf(g0, wf, af0[j]);
            af1[j] = fmaf(g1, wf, af1[j]);
            ac0[j] = fmaf(g0, wc, ac0[j]);
            ac1[j] = fmaf(g1, wc, ac1[j]);
        }
    }
    __device__ __forceinline__ void finish(float* __restrict__ slab) {
#pragma unroll
        for (int j = 0; j < 4; ++j) flush(slab, j);
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// CellRun: the scatter walk with ONE run test per sample instead of one per corner pair.  The eight target rows of a sample are
// rf[j] (x-floor) and rf[j] ^ dxr (x-ceil), j = (y,z) combination, dxr = (ix_floor ^ ix_ceil) & mask (the hash is linear in x).
// While the integer cell (floor and ceil coordinates) of consecutive samples stays the same the sixteen partial sums stay in
// registers; a change flushes all four pairs at once: dxr == 1 -> one 16-byte red.v4 per pair (rows r, r^1), dxr == 0 (x exactly on
// a grid plane: both x rows coincide and the x-ceil weight is 0) -> one red.v2 per pair, else two.  ~75 instructions per
// (sample, level) against ~300 for ScatterRun's per-pair tests and flushes.
// ---------------------------------------------------------------------------------------------------------------------
struct CellRun {
    int kf[3], kc[3];  // integer cell of the pending run (floor / ceil coordinates); kf[0] < 0: nothing pending
    uint32_t rf[4], dxr;
    float af0[4], af1[4], ac0[4], ac1[4];
    __device__ __forceinline__ void reset() { kf[0] = -1; }
    __device__ __forceinline__ void flush(float* __restrict__ slab) {
        if (kf[0] < 0) return;
        if (dxr == 1u) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool odd = rf[j] & 1u;  // x-floor row is the upper member of the aligned pair
                nvo_red_add_v4(slab + 2 * (size_t)(rf[j] & ~1u), odd ? ac0[j] : af0[j], odd ? ac1[j] : af1[j], odd ? af0[j] : ac0[j], odd ? af1[j] : ac1[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) nvo_red_add_v2(slab + 2 * (size_t)rf[j], af0[j], af1[j]);
            if (dxr != 0u) {
#pragma unroll
                for (int j = 0; j < 4; ++j) nvo_red_add_v2(slab + 2 * (size_t)(rf[j] ^ dxr), ac0[j], ac1[j]);
            }
        }
    }
    // add one sample's (g0, g1) at normalised position (px, py, pz) of the level with `scale`
    __device__ __forceinline__ void add(float* __restrict__ slab, float px, float py, float pz, float scale, uint32_t mask, float g0, float g1) {
        const float sx = __fmul_rn(px, scale), sy = __fmul_rn(py, scale), sz = __fmul_rn(pz, scale);
        const float fx = floorf(sx), fy = floorf(sy), fz = floorf(sz);
        const float ox = sx - fx, oy = sy - fy, oz = sz - fz;
        const int ifx = (int)fx, ify = (int)fy, ifz = (int)fz;
        const int icx = ifx + (sx != fx), icy = ify + (sy != fy), icz = ifz + (sz != fz);
        if (((ifx ^ kf[0]) | (ify ^ kf[1]) | (ifz ^ kf[2]) | (icx ^ kc[0]) | (icy ^ kc[1]) | (icz ^ kc[2])) != 0) {
            flush(slab);
            kf[0] = ifx, kf[1] = ify, kf[2] = ifz, kc[0] = icx, kc[1] = icy, kc[2] = icz;
            const uint32_t hy0 = (uint32_t)ify * PRIME_Y, hy1 = (uint32_t)icy * PRIME_Y, hz0 = (uint32_t)ifz * PRIME_Z, hz1 = (uint32_t)icz * PRIME_Z;
            rf[0] = ((uint32_t)ifx ^ hy0 ^ hz0) & mask;  // j = sy + 2*sz
            rf[1] = ((uint32_t)ifx ^ hy1 ^ hz0) & mask;
            rf[2] = ((uint32_t)ifx ^ hy0 ^ hz1) & mask;
            rf[3] = ((uint32_t)ifx ^ hy1 ^ hz1) & mask;
            dxr = ((uint32_t)ifx ^ (uint32_t)icx) & mask;
#pragma unroll
            for (int j = 0; j < 4; ++j) af0[j] = af1[j] = ac0[j] = ac1[j] = 0.f;
        }
        const float wy0 = 1.f - oy, wz0 = 1.f - oz;
        const float wyz[4] = {wz0 * wy0, wz0 * oy, oz * wy0, oz * oy};
        const float wx0 = 1.f - ox;
        const float gf0 = g0 * wx0, gf1 = g1 * wx0, gc0 = g0 * ox, gc1 = g1 * ox;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            af0[j] = fmaf(gf0, wyz[j], af0[j]);
            af1[j] = fmaf(gf1, wyz[j], af1[j]);
            ac0[j] = fmaf(gc0, wyz[j], ac0[j]);
            ac1[j] = fmaf(gc1, wyz[j], ac1[j]);
        }
    }
};
