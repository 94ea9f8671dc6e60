// The nerfacto field's networks as ONE persistent tcgen05 kernel per direction (NS/fields/nerfacto_field.py:199-297):
//   hash features -> mlp_base (32->64->16) -> density, geometry features
//                 -> density-gradient normals (base_field.py:80-101: the base network's input-gradient chain x the saved d feature / dx)
//                 -> input assembly (SH16 | geo15 | appearance32, posenc12 | geo15) -> mlp_head (63->64->64->3, sigmoid)
//                                                                                      -> mlp_pred_normals + head (27->64->64->64->3, tanh, normalize)
// Nothing between the hash features and the per-sample outputs goes through HBM except what the backward needs (saved activations).
//
// Structure (csrc/mlp_tc.cu evaluated the same three networks as five launches with a serial wait -> MMA -> epilogue -> barrier chain each):
//   * one CTA per SM, G "tile groups" of 256 threads; each group walks its own 128-sample tiles through the 8-step layer chain with its own
//     TMEM accumulator columns (128 per group), activation buffers, mbarriers and named barrier, so while one group waits for its MMA or runs
//     an epilogue the tensor pipe executes another group's layer: the groups' chains overlap, the weights are resident once;
//   * all weights (60 KB of fp16 tiles) are pulled into shared memory by one bulk copy per CTA; biases ride in the MMA: every layer gets one
//     extra K = 16 step whose A operand is a constant "ones" tile and whose B operand holds the bias (the head / pred-normals input layers
//     use a spare input column instead), so the epilogue of a hidden layer is tcgen05.ld -> cvt.rn.relu.f16x2 -> 16-byte stores;
//   * the two narrow output layers (N = 16) are issued together with the next wide layer (separate accumulator columns, one commit).
// Weight images are K-major "natural tiles": element (n, k) of a layer with N outputs at byte ((k >> 3) * N + n) * 16 + (k & 7) * 2.
#include "nvo_common.cuh"
#include "tc_common.cuh"

#define FT_THREADS 256       // per tile group: warp w reads TMEM lanes 32 * (w & 3) .. (rows of the tile) and the column half (w >> 2)
#define GEO 15
#define APP 32

// ---- weight image -------------------------------------------------------------------------------------------------------------------
// section byte offsets; "kd" = data K (multiple of 16), bias step appended where `bias`
#define FI_B0 0                                   // N 64, kd 32 + bias         6144
#define FI_B1 (FI_B0 + 64 * 48 * 2)               // N 16, kd 64 + bias         2560
#define FI_NS (FI_B1 + 16 * 80 * 2)               // N 32, kd 64 (W0 x level scale, transposed use)  4096
#define FI_H0 (FI_NS + 32 * 64 * 2)               // N 64, kd 64 (col 63 = bias) 8192
#define FI_H1 (FI_H0 + 64 * 64 * 2)               // N 64, kd 64 + bias         10240
#define FI_H2 (FI_H1 + 64 * 80 * 2)               // N 16, kd 64 + bias         2560
#define FI_P0 (FI_H2 + 16 * 80 * 2)               // N 64, kd 32 (col 27 = bias) 4096
#define FI_P1 (FI_P0 + 64 * 32 * 2)               // N 64, kd 64 + bias         10240
#define FI_P2 (FI_P1 + 64 * 80 * 2)               // N 64, kd 64 + bias         10240
#define FI_P3 (FI_P2 + 64 * 80 * 2)               // N 16, kd 64 + bias         2560
#define FI_W1ROW (FI_P3 + 16 * 80 * 2)            // fp16 [64]: W_base1[0][:], the raw-density row (seed of the normals chain)  128
#define FI_BYTES (FI_W1ROW + 128)

// saved-activation tile (backward): chunk offsets inside a tile of FS_CHUNKS 2 KB chunks
#define FS_H1 0     // mlp_base hidden (64)
#define FS_X 8      // head input (64)
#define FS_P 16     // pred-normals input (32)
#define FS_AH1 20
#define FS_AH2 28
#define FS_AP1 36
#define FS_AP2 44
#define FS_AP3 52
#define FS_CHUNKS 60

struct FieldFwdP {
    int64_t n;
    int S;
    int want_pn, want_normals;
    const unsigned char* feat16;  // TMH [tiles][4][128][8] fp16
    const uint4* jac;             // [tiles][4][3][128] x 16 B
    const float* pos;             // [n,3] sample positions (world)
    const float* dirs;            // [B,3]
    const int64_t* cam;           // [B] or null (eval: `emb` is one 32-vector)
    const float* emb;
    const float* sel;             // [n]
    const unsigned char* wimg;
    float* density;               // [n]
    float* rgb;                   // [n,3]
    float* pn;                    // [n,3] normalised predicted normals
    float* normals;               // [n,3]
    float* h0;                    // [n] raw density (saved for trunc_exp')
    float* pn_raw;                // [n,3] tanh output before normalisation (saved)
    uint4* saved;                 // [tiles][FS_CHUNKS][128] x 16 B or null
};

// ---- small device helpers -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cvt_h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t cvt_relu_h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint4 pack8f(const float* v) { return make_uint4(cvt_h2(v[0], v[1]), cvt_h2(v[2], v[3]), cvt_h2(v[4], v[5]), cvt_h2(v[6], v[7])); }
__device__ __forceinline__ void group_sync(int g) {
    // generic-proxy writes of the epilogue (shared memory operands) and its TMEM reads are ordered before the MMAs the group leader issues next
    tc_fence_before();
    fence_async_smem();
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(FT_THREADS) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// one layer's MMAs: D[128 x N] (tmem_d) = A (ksteps x 16 columns from a_base) x W^T (+ bias step against the ones tile)
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_base, int ksteps, uint32_t w_base, int N, bool bias, uint32_t ones_base) {
    const uint32_t idesc = umma_idesc(TM, N, 0, 0);
    for (int k = 0; k < ksteps; ++k)
        umma_f16(tmem_d, umma_desc(a_base + k * 2 * CHUNK_B, CHUNK_B, 128), umma_desc(w_base + k * 2 * N * 16, N * 16, 128), idesc, k > 0);
    if (bias) umma_f16(tmem_d, umma_desc(ones_base, CHUNK_B, 128), umma_desc(w_base + ksteps * 2 * N * 16, N * 16, 128), idesc, 1);
}

// hidden-layer epilogue: this warp's 32 accumulator columns -> (ReLU) -> fp16 -> the next layer's A tile (+ the saved tile)
template <bool RELU>
__device__ __forceinline__ void epi_hidden(uint32_t trow, unsigned char* __restrict__ sOut, uint4* __restrict__ gsave, int hf, int r) {
    float v[32];
    tmem_ld32(trow + 32 * hf, v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint4 u;
        if (RELU)
            u = make_uint4(cvt_relu_h2(v[8 * q], v[8 * q + 1]), cvt_relu_h2(v[8 * q + 2], v[8 * q + 3]), cvt_relu_h2(v[8 * q + 4], v[8 * q + 5]),
                           cvt_relu_h2(v[8 * q + 6], v[8 * q + 7]));
        else
            u = pack8f(v + 8 * q);
        *reinterpret_cast<uint4*>(sOut + (4 * hf + q) * CHUNK_B + r * 16) = u;
        if (gsave) gsave[(4 * hf + q) * TM + r] = u;
    }
}

// SH degree 4 and the torch-path frequency encoding, same arithmetic as csrc/field.cu (NS/utils/math.py:45-78, encodings.py:170-176)
__device__ __forceinline__ void ft_sh16(float x, float y, float z, float* c) {
    const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    c[0] = 0.28209479177387814f;
    c[1] = __fmul_rn(0.4886025119029199f, y);
    c[2] = __fmul_rn(0.4886025119029199f, z);
    c[3] = __fmul_rn(0.4886025119029199f, x);
    c[4] = __fmul_rn(__fmul_rn(1.0925484305920792f, x), y);
    c[5] = __fmul_rn(__fmul_rn(1.0925484305920792f, y), z);
    c[6] = __fsub_rn(__fmul_rn(0.9461746957575601f, zz), 0.31539156525251999f);
    c[7] = __fmul_rn(__fmul_rn(1.0925484305920792f, x), z);
    c[8] = __fmul_rn(0.5462742152960396f, __fsub_rn(xx, yy));
    c[9] = __fmul_rn(__fmul_rn(0.5900435899266435f, y), __fsub_rn(__fmul_rn(3.f, xx), yy));
    c[10] = __fmul_rn(__fmul_rn(__fmul_rn(2.890611442640554f, x), y), z);
    c[11] = __fmul_rn(__fmul_rn(0.4570457994644658f, y), __fsub_rn(__fmul_rn(5.f, zz), 1.f));
    c[12] = __fmul_rn(__fmul_rn(0.3731763325901154f, z), __fsub_rn(__fmul_rn(5.f, zz), 3.f));
    c[13] = __fmul_rn(__fmul_rn(0.4570457994644658f, x), __fsub_rn(__fmul_rn(5.f, zz), 1.f));
    c[14] = __fmul_rn(__fmul_rn(1.445305721320277f, z), __fsub_rn(xx, yy));
    c[15] = __fmul_rn(__fmul_rn(0.5900435899266435f, x), __fsub_rn(xx, __fmul_rn(3.f, yy)));
}
__device__ __forceinline__ float ft_posenc(float xi, int k, bool cos_block) {
    const float u = __fmul_rn(__fmul_rn(6.283185307179586f, xi), (float)(1 << k));
    return sinf(cos_block ? __fadd_rn(u, 1.5707963267948966f) : u);
}

// ================================================================================================================================
// forward
// shared memory: [weight image FI_BYTES][ones tile 2 chunks][per group: FI 4 | H 8 | X 8 | P 4 chunks][mbarriers][tmem ptr]
// TMEM per group: accA = columns [0,64) (wide layers; normals chain in [32,64) next to the 16-wide base output), accB = [64,80) (the
// narrow output layers issued together with the next wide one)
// ================================================================================================================================
#define FT_GROUP_CHUNKS 24
#define FT_ONES_OFF ((FI_BYTES + 1023) & ~1023)
#define FT_GROUPS_OFF (FT_ONES_OFF + 2 * CHUNK_B)

template <int G>
__global__ void __launch_bounds__(G* FT_THREADS, 1) k_field_fwd(const __grid_constant__ FieldFwdP p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int g = threadIdx.x >> 8, tid = threadIdx.x & 255, warp = tid >> 5, hf = warp >> 2;
    const int r = ((warp & 3) << 5) | (tid & 31);  // row of the tile == TMEM lane
    unsigned char* sW = smem;
    unsigned char* sOnes = smem + FT_ONES_OFF;
    unsigned char* sG = smem + FT_GROUPS_OFF + g * FT_GROUP_CHUNKS * CHUNK_B;
    unsigned char *sFI = sG, *sH = sG + 4 * CHUNK_B, *sX = sG + 12 * CHUNK_B, *sP = sG + 20 * CHUNK_B;
    uint64_t* mbars = reinterpret_cast<uint64_t*>(smem + FT_GROUPS_OFF + G * FT_GROUP_CHUNKS * CHUNK_B);
    uint64_t* mbar_w = mbars;
    uint64_t* mbar_mma = mbars + 1 + g;
    uint64_t* mbar_in = mbars + 1 + G + g;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(mbars + 1 + 2 * G);
    constexpr int TMEM_COLS = G * 128 <= 256 ? 256 : 512;
    const int64_t n_tiles = (p.n + TM - 1) >> 7;
    const int64_t stride = (int64_t)gridDim.x * G;
    const int64_t first = (int64_t)blockIdx.x * G + g;

    // ones tile: chunk 0 = feature 0 is 1.0 for every row, chunk 1 = zeros
    for (int e = threadIdx.x; e < 2 * CHUNK_B / 16; e += G * FT_THREADS)
        reinterpret_cast<uint4*>(sOnes)[e] = e < TM ? make_uint4(0x00003C00u, 0, 0, 0) : make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 1 + 2 * G; ++i) mbar_init(mbars + i, 1);
        fence_mbar_init();
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        mbar_expect_tx(mbar_w, FI_BYTES);
        bulk_g2s(sW, p.wimg, FI_BYTES, mbar_w);
    }
    if (tid == 0 && first < n_tiles) {
        mbar_expect_tx(mbar_in, 4 * CHUNK_B);
        bulk_g2s(sFI, p.feat16 + first * (4 * CHUNK_B), 4 * CHUNK_B, mbar_in);
    }
    const uint32_t tmem = *tmem_ptr + (uint32_t)(g * 128);
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's lanes
    const uint32_t accA = tmem, accB = tmem + 64;
    const uint32_t uW = smem_u32(sW), uOnes = smem_u32(sOnes), uFI = smem_u32(sFI), uH = smem_u32(sH), uX = smem_u32(sX), uP = smem_u32(sP);
    mbar_wait(mbar_w, 0);
    uint32_t ph = 0;
    int it = 0;
#define FT_COMMIT_WAIT()                   \
    mbar_wait(mbar_mma, ph);               \
    ph ^= 1;                               \
    tc_fence_after();

    for (int64_t tile = first; tile < n_tiles; tile += stride, ++it) {
        const int64_t t = tile * TM + r;
        const bool live = t < p.n;
        const int64_t tc = live ? t : p.n - 1;  // rows past the end read the last sample's inputs (finite) and write nothing
        uint4* sv = p.saved ? p.saved + tile * (int64_t)(FS_CHUNKS * TM) : nullptr;
        mbar_wait(mbar_in, (uint32_t)(it & 1));
        // ---- S0: mlp_base layer 0 --------------------------------------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            issue_layer(accA, uFI, 2, uW + FI_B0, 64, true, uOnes);
            umma_commit(mbar_mma);
        }
        FT_COMMIT_WAIT();
        if (tid == 0 && tile + stride < n_tiles) {  // the feature tile has been consumed: fetch the next one behind the rest of the chain
            mbar_expect_tx(mbar_in, 4 * CHUNK_B);
            bulk_g2s(sFI, p.feat16 + (tile + stride) * (4 * CHUNK_B), 4 * CHUNK_B, mbar_in);
        }
        {
            float v[32];
            tmem_ld32(trow + 32 * hf, v);
            const uint4* w1 = reinterpret_cast<const uint4*>(sW + FI_W1ROW) + 4 * hf;  // W_base1[0][32 hf ..]: warp-broadcast reads
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 u = make_uint4(cvt_relu_h2(v[8 * q], v[8 * q + 1]), cvt_relu_h2(v[8 * q + 2], v[8 * q + 3]), cvt_relu_h2(v[8 * q + 4], v[8 * q + 5]),
                                           cvt_relu_h2(v[8 * q + 6], v[8 * q + 7]));
                *reinterpret_cast<uint4*>(sH + (4 * hf + q) * CHUNK_B + r * 16) = u;
                if (sv) sv[(FS_H1 + 4 * hf + q) * TM + r] = u;
                if (p.want_normals) {
                    // dZ of the hidden layer for d raw_density / d features: ReLU'(a) * W1[0][j]  (base_field.py:92-97, grad_outputs = 1)
                    const uint4 w = w1[q];
                    const __half2 zero = __float2half2_rn(0.f);
                    uint4 d;
                    const __half2* a2 = reinterpret_cast<const __half2*>(&u);
                    const __half2* w2 = reinterpret_cast<const __half2*>(&w);
                    __half2* d2 = reinterpret_cast<__half2*>(&d);
#pragma unroll
                    for (int j = 0; j < 4; ++j) d2[j] = __hmul2(__hgt2(a2[j], zero), w2[j]);
                    *reinterpret_cast<uint4*>(sX + (4 * hf + q) * CHUNK_B + r * 16) = d;
                }
            }
        }
        group_sync(g);
        // ---- S1: mlp_base layer 1 (16 outputs) + the normals chain's input-gradient product (32 feature columns) -------------------------
        if (tid == 0) {
            tc_fence_after();
            issue_layer(accA, uH, 4, uW + FI_B1, 16, true, uOnes);
            if (p.want_normals) issue_layer(accA + 32, uX, 4, uW + FI_NS, 32, false, uOnes);
            umma_commit(mbar_mma);
        }
        FT_COMMIT_WAIT();
        {
            const int64_t ray = tc / p.S;
            float hv[16];
            tmem_ld16(trow, hv);  // both halves need the geometry features
            if (hf == 0) {
                // density = trunc_exp(h0) * selector (nerfacto_field.py:216-221), head input columns 0..31, pred-normals input
                if (live) {
                    p.density[t] = __fmul_rn(expf(hv[0]), __ldg(p.sel + t));
                    if (p.h0) p.h0[t] = hv[0];
                }
                float c[16];
                ft_sh16(__fmul_rn(__fadd_rn(__ldg(p.dirs + 3 * ray), 1.f), 0.5f), __fmul_rn(__fadd_rn(__ldg(p.dirs + 3 * ray + 1), 1.f), 0.5f),
                        __fmul_rn(__fadd_rn(__ldg(p.dirs + 3 * ray + 2), 1.f), 0.5f), c);
                const float e0 = __ldg((p.cam ? p.emb + APP * __ldg(p.cam + ray) : p.emb));
                uint4 x0 = pack8f(c), x1 = pack8f(c + 8), x2 = pack8f(hv + 1);
                uint4 x3 = make_uint4(cvt_h2(hv[9], hv[10]), cvt_h2(hv[11], hv[12]), cvt_h2(hv[13], hv[14]), cvt_h2(hv[15], e0));
                *reinterpret_cast<uint4*>(sX + 0 * CHUNK_B + r * 16) = x0;
                *reinterpret_cast<uint4*>(sX + 1 * CHUNK_B + r * 16) = x1;
                *reinterpret_cast<uint4*>(sX + 2 * CHUNK_B + r * 16) = x2;
                *reinterpret_cast<uint4*>(sX + 3 * CHUNK_B + r * 16) = x3;
                if (sv) sv[(FS_X + 0) * TM + r] = x0, sv[(FS_X + 1) * TM + r] = x1, sv[(FS_X + 2) * TM + r] = x2, sv[(FS_X + 3) * TM + r] = x3;
                if (p.want_pn) {
                    float pe[12];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const float xi = __ldg(p.pos + 3 * tc + i);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            pe[i * 2 + k] = ft_posenc(xi, k, false);
                            pe[6 + i * 2 + k] = ft_posenc(xi, k, true);
                        }
                    }
                    const uint4 p0 = pack8f(pe);
                    const uint4 p1 = make_uint4(cvt_h2(pe[8], pe[9]), cvt_h2(pe[10], pe[11]), cvt_h2(hv[1], hv[2]), cvt_h2(hv[3], hv[4]));
                    const uint4 p2 = pack8f(hv + 5);
                    const uint4 p3 = make_uint4(cvt_h2(hv[13], hv[14]), cvt_h2(hv[15], 1.f), 0u, 0u);  // column 27 = 1: carries the first layer's bias
                    *reinterpret_cast<uint4*>(sP + 0 * CHUNK_B + r * 16) = p0;
                    *reinterpret_cast<uint4*>(sP + 1 * CHUNK_B + r * 16) = p1;
                    *reinterpret_cast<uint4*>(sP + 2 * CHUNK_B + r * 16) = p2;
                    *reinterpret_cast<uint4*>(sP + 3 * CHUNK_B + r * 16) = p3;
                    if (sv) sv[(FS_P + 0) * TM + r] = p0, sv[(FS_P + 1) * TM + r] = p1, sv[(FS_P + 2) * TM + r] = p2, sv[(FS_P + 3) * TM + r] = p3;
                }
            } else {
                // density-gradient normals first (they read accA[32,64) and the buffer the head input is about to overwrite was its operand)
                if (p.want_normals) {
                    float v[32];
                    tmem_ld32(trow + 32, v);
                    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint4* jb = p.jac + ((tile * 4 + c) * 3) * TM + r;
                        const uint4 J[3] = {__ldg(jb), __ldg(jb + TM), __ldg(jb + 2 * TM)};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float g0 = v[8 * c + 2 * q], g1 = v[8 * c + 2 * q + 1];  // already x level scale (folded into the weights)
#pragma unroll
                            for (int a = 0; a < 3; ++a) {
                                const uint32_t w = q == 0 ? J[a].x : q == 1 ? J[a].y : q == 2 ? J[a].z : J[a].w;
                                const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&w));
                                acc[a] = fmaf(d.x, g0, fmaf(d.y, g1, acc[a]));
                            }
                        }
                    }
                    const float nrm = fmaxf(sqrtf(acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2]), 1e-12f);
                    if (live) {
                        p.normals[3 * t] = -(acc[0] / nrm);
                        p.normals[3 * t + 1] = -(acc[1] / nrm);
                        p.normals[3 * t + 2] = -(acc[2] / nrm);
                    }
                }
                // head input columns 32..63: appearance embedding 1..31, then the constant 1 that carries the first layer's bias
                const float* e = p.cam ? p.emb + APP * __ldg(p.cam + ray) : p.emb;
                float ev[32];
#pragma unroll
                for (int k = 0; k < APP; k += 4) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(e + k));
                    ev[k] = q.x, ev[k + 1] = q.y, ev[k + 2] = q.z, ev[k + 3] = q.w;
                }
                const uint4 x4 = pack8f(ev + 1), x5 = pack8f(ev + 9), x6 = pack8f(ev + 17);
                const uint4 x7 = make_uint4(cvt_h2(ev[25], ev[26]), cvt_h2(ev[27], ev[28]), cvt_h2(ev[29], ev[30]), cvt_h2(ev[31], 1.f));
                // (the normals MMA read sX as its operand: it retired before the commit this epilogue waited for)
                *reinterpret_cast<uint4*>(sX + 4 * CHUNK_B + r * 16) = x4;
                *reinterpret_cast<uint4*>(sX + 5 * CHUNK_B + r * 16) = x5;
                *reinterpret_cast<uint4*>(sX + 6 * CHUNK_B + r * 16) = x6;
                *reinterpret_cast<uint4*>(sX + 7 * CHUNK_B + r * 16) = x7;
                if (sv) sv[(FS_X + 4) * TM + r] = x4, sv[(FS_X + 5) * TM + r] = x5, sv[(FS_X + 6) * TM + r] = x6, sv[(FS_X + 7) * TM + r] = x7;
            }
        }
        group_sync(g);
        // ---- S2 / S3: mlp_head hidden layers ------------------------------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            issue_layer(accA, uX, 4, uW + FI_H0, 64, false, uOnes);
            umma_commit(mbar_mma);
        }
        FT_COMMIT_WAIT();
        epi_hidden<true>(trow, sH, sv ? sv + FS_AH1 * TM : nullptr, hf, r);
        group_sync(g);
        if (tid == 0) {
            tc_fence_after();
            issue_layer(accA, uH, 4, uW + FI_H1, 64, true, uOnes);
            umma_commit(mbar_mma);
        }
        FT_COMMIT_WAIT();
        epi_hidden<true>(trow, sH, sv ? sv + FS_AH2 * TM : nullptr, hf, r);
        group_sync(g);
        // ---- S4 (+ S5): colour output layer, issued together with the pred-normals input layer ---------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            issue_layer(accB, uH, 4, uW + FI_H2, 16, true, uOnes);
            if (p.want_pn) issue_layer(accA, uP, 2, uW + FI_P0, 64, false, uOnes);
            umma_commit(mbar_mma);
        }
        FT_COMMIT_WAIT();
        if (hf == 0) {
            float v[4];
            tmem_ld4(trow + 64, v);
            if (live) {
#pragma unroll
                for (int j = 0; j < 3; ++j) p.rgb[3 * t + j] = 1.f / (1.f + expf(-v[j]));  // Sigmoid (nerfacto_field.py:196)
            }
        }
        if (p.want_pn) {
            epi_hidden<true>(trow, sH, sv ? sv + FS_AP1 * TM : nullptr, hf, r);
            group_sync(g);
            // ---- S6 / S7: mlp_pred_normals layers 1, 2 (the last one has no activation, mlp.py:143-179 with out_activation None) -------------
            if (tid == 0) {
                tc_fence_after();
                issue_layer(accA, uH, 4, uW + FI_P1, 64, true, uOnes);
                umma_commit(mbar_mma);
            }
            FT_COMMIT_WAIT();
            epi_hidden<true>(trow, sH, sv ? sv + FS_AP2 * TM : nullptr, hf, r);
            group_sync(g);
            if (tid == 0) {
                tc_fence_after();
                issue_layer(accA, uH, 4, uW + FI_P2, 64, true, uOnes);
                umma_commit(mbar_mma);
            }
            FT_COMMIT_WAIT();
            epi_hidden<false>(trow, sH, sv ? sv + FS_AP3 * TM : nullptr, hf, r);
            group_sync(g);
            // ---- S8: PredNormalsFieldHead: Linear(64, 3) + Tanh + normalize (field_heads.py:189-204) -------------------------------------------
            if (tid == 0) {
                tc_fence_after();
                issue_layer(accB, uH, 4, uW + FI_P3, 16, true, uOnes);
                umma_commit(mbar_mma);
            }
            FT_COMMIT_WAIT();
            if (hf == 0) {
                float v[4];
                tmem_ld4(trow + 64, v);
                if (live) {
                    const float a = tanhf(v[0]), b = tanhf(v[1]), c = tanhf(v[2]);
                    if (p.pn_raw) p.pn_raw[3 * t] = a, p.pn_raw[3 * t + 1] = b, p.pn_raw[3 * t + 2] = c;
                    const float nrm = fmaxf(sqrtf(a * a + b * b + c * c), 1e-12f);
                    p.pn[3 * t] = a / nrm, p.pn[3 * t + 1] = b / nrm, p.pn[3 * t + 2] = c / nrm;
                }
            }
        }
        group_sync(g);  // the next tile's first MMA overwrites accA and (after its epilogue) sH
    }
#undef FT_COMMIT_WAIT
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_ptr), "n"(TMEM_COLS) : "memory");
}

// ================================================================================================================================
// weight image packing: torch-layout fp32 parameters ([W0 [out,in], b0, W1, b1, ...] per network) -> the fp16 tiles above
// ================================================================================================================================
struct GridScales {
    float s[16];
};

struct PackLayer {
    int src;            // 0 base, 1 head, 2 pred-normals
    int w_off, b_off;   // float offsets inside the network's flat parameters
    int N, K;           // real outputs / inputs
    int np, kd;         // padded outputs, data K of the tile
    int bias_col;       // column that carries the bias (kd = bias step; < kd = spare input column)
    int img_off;
    int ktot;           // kd (+16 with a bias step)
};

__global__ void __launch_bounds__(256) k_field_pack(const float* __restrict__ base, const float* __restrict__ head, const float* __restrict__ pnp,
                                                    const __grid_constant__ GridScales sc, unsigned char* __restrict__ img) {
    // layer table (nerfacto_field.py:130-197): base 32->64->16, head 63->64->64->3, pred-normals 27->64->64->64 (+ Linear 64->3)
    const PackLayer L[9] = {
        {0, 0, 64 * 32, 64, 32, 64, 32, 32, FI_B0, 48},
        {0, 64 * 32 + 64, 64 * 32 + 64 + 16 * 64, 16, 64, 16, 64, 64, FI_B1, 80},
        {1, 0, 64 * 63, 64, 63, 64, 64, 63, FI_H0, 64},
        {1, 64 * 63 + 64, 64 * 63 + 64 + 64 * 64, 64, 64, 64, 64, 64, FI_H1, 80},
        {1, 64 * 63 + 64 + 64 * 64 + 64, 64 * 63 + 64 + 64 * 64 + 64 + 3 * 64, 3, 64, 16, 64, 64, FI_H2, 80},
        {2, 0, 64 * 27, 64, 27, 64, 32, 27, FI_P0, 32},
        {2, 64 * 27 + 64, 64 * 27 + 64 + 64 * 64, 64, 64, 64, 64, 64, FI_P1, 80},
        {2, 64 * 27 + 64 + 64 * 64 + 64, 64 * 27 + 64 + 64 * 64 + 64 + 64 * 64, 64, 64, 64, 64, 64, FI_P2, 80},
        {2, 64 * 27 + 64 + 2 * (64 * 64 + 64), 64 * 27 + 64 + 2 * (64 * 64 + 64) + 3 * 64, 3, 64, 16, 64, 64, FI_P3, 80},
    };
    const int t = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int l = 0; l < 9; ++l) {
        const PackLayer& y = L[l];
        const float* src = y.src == 0 ? base : (y.src == 1 ? head : pnp);
        if (!src) continue;
        __half* iw = reinterpret_cast<__half*>(img + y.img_off);
        for (int e = t; e < y.np * y.ktot; e += stride) {
            const int o = e / y.ktot, i = e - o * y.ktot;
            float w = 0.f;
            if (o < y.N) {
                if (i < y.K)
                    w = __ldg(src + y.w_off + o * y.K + i);
                else if (i == y.bias_col)
                    w = __ldg(src + y.b_off + o);
            }
            iw[((i >> 3) * y.np + o) * 8 + (i & 7)] = __float2half_rn(w);
        }
    }
    // normals chain: Ws[i][o] = W_base0[o][i] * scale_{i/2} as a tile with N = 32 (features i), K = 64 (hidden units o)
    __half* ns = reinterpret_cast<__half*>(img + FI_NS);
    for (int e = t; e < 32 * 64; e += stride) {
        const int i = e >> 6, o = e & 63;
        ns[((o >> 3) * 32 + i) * 8 + (o & 7)] = __float2half_rn(__ldg(base + o * 32 + i) * sc.s[i >> 1]);
    }
    __half* w1 = reinterpret_cast<__half*>(img + FI_W1ROW);
    for (int o = t; o < 64; o += stride) w1[o] = __float2half_rn(__ldg(base + 64 * 32 + 64 + o));
}

// ================================================================================================================================
extern "C" int64_t nvo_field_wimage_bytes(void) { return FI_BYTES; }
extern "C" int64_t nvo_field_saved_bytes(int64_t n) { return ((n + TM - 1) / TM) * (int64_t)FS_CHUNKS * CHUNK_B; }

extern "C" int nvo_field_pack_weights(void* stream, const nvo_grid_desc* grid, const float* base_params, const float* head_params, const float* pn_params,
                                      void* wimage) {
    NVO_CHECK(grid && grid->n_levels == 16, "field_pack_weights: the fused field kernels are specialised for the 16-level main grid");
    NVO_CHECK(base_params && head_params && wimage, "field_pack_weights: null pointer");
    GridScales sc;
    for (int i = 0; i < 16; ++i) sc.s[i] = grid->scalings[i];
    k_field_pack<<<16, 256, 0, (cudaStream_t)stream>>>(base_params, head_params, pn_params, sc, (unsigned char*)wimage);
    NVO_CUDA_LAUNCH_CHECK("field_pack_weights");
    return 0;
}

static int field_groups() {
    static const int g = nvo_env_int("NVO_FIELD_GROUPS", 3);
    return g == 2 ? 2 : 3;
}

extern "C" int nvo_field_forward(void* stream, int64_t B, int32_t S, const void* feat16, const void* jac, const float* positions, const float* directions,
                                 const int64_t* cam_idx, const float* embedding, const float* selector, const void* wimage, float* density, float* rgb,
                                 float* pred_normals, float* normals, float* h0, float* pn_raw, void* saved) {
    NVO_CHECK(B >= 0 && S >= 1, "field_forward: bad shape B=%lld S=%d", (long long)B, S);
    if (B == 0) return 0;
    NVO_CHECK(feat16 && directions && embedding && selector && wimage && density && rgb, "field_forward: null pointer");
    NVO_CHECK(!pred_normals || positions, "field_forward: positions required for the predicted normals");
    NVO_CHECK(!normals || jac, "field_forward: the saved feature derivatives (nvo_grid_forward_jac) are required for the normals");
    FieldFwdP p;
    p.n = B * S, p.S = S, p.want_pn = pred_normals != nullptr, p.want_normals = normals != nullptr;
    p.feat16 = (const unsigned char*)feat16, p.jac = (const uint4*)jac, p.pos = positions, p.dirs = directions, p.cam = cam_idx, p.emb = embedding;
    p.sel = selector, p.wimg = (const unsigned char*)wimage, p.density = density, p.rgb = rgb, p.pn = pred_normals, p.normals = normals, p.h0 = h0;
    p.pn_raw = pn_raw, p.saved = (uint4*)saved;
    const int G = field_groups();
    const size_t smem = FT_GROUPS_OFF + (size_t)G * FT_GROUP_CHUNKS * CHUNK_B + 8 * (1 + 2 * G) + 16;
    const int64_t tiles = (p.n + TM - 1) / TM;
    const unsigned int grid = (unsigned int)min((int64_t)nvo_sm_count(), (tiles + G - 1) / G);
    cudaError_t e;
    if (G == 2) {
        e = cudaFuncSetAttribute(k_field_fwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        NVO_CHECK(e == cudaSuccess, "field_forward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        k_field_fwd<2><<<grid, 2 * FT_THREADS, smem, (cudaStream_t)stream>>>(p);
    } else {
        e = cudaFuncSetAttribute(k_field_fwd<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        NVO_CHECK(e == cudaSuccess, "field_forward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        k_field_fwd<3><<<grid, 3 * FT_THREADS, smem, (cudaStream_t)stream>>>(p);
    }
    NVO_CUDA_LAUNCH_CHECK("field_forward");
    return 0;
}
