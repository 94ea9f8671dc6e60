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

// phase timestamps of one tile group (tools/field_timing.cu builds this file with -DNVO_FT_TIMING); compiled out of the library
#ifdef NVO_FT_TIMING
__device__ unsigned long long* g_ft_timing = nullptr;  // [2 threads][FT_MAX_MARKS]
#define FT_MAX_MARKS 256
#define FT_MARK()                                                                                              \
    do {                                                                                                       \
        if (blockIdx.x == 0 && g == 0 && (tid == 0 || tid == 128) && g_ft_timing && ft_mark < FT_MAX_MARKS)    \
            g_ft_timing[(tid >> 7) * FT_MAX_MARKS + ft_mark++] = clock64();                                    \
    } while (0)
// backward: [0] = group 0's thread 0, [1] = the issuer's lane 0 (events of group 0), CTA 0
__device__ unsigned long long* g_fb_timing = nullptr;  // [2][FT_MAX_MARKS]
__device__ unsigned long long* g_fb_cta = nullptr;  // [gridDim.x][8] %globaltimer (ns): kernel entry, setup done, tile loop done, weight gradients flushed, then thread 0's flush: first TMEM read done, first stores issued, last stores issued
__device__ __forceinline__ unsigned long long fb_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define FB_CTA_MARK(i)                                                                     \
    do {                                                                                   \
        if (threadIdx.x == 0 && g_fb_cta) g_fb_cta[8 * blockIdx.x + (i)] = fb_gtime();     \
    } while (0)
__device__ int g_fb_skip = 0;  // timing experiments only (results are wrong): bit 0 = no wgrad MMAs, bit 1 = no dgrad MMAs, bit 2 = no slot copies after the first two
#define FB_SKIP(bit) ((g_fb_skip >> (bit)) & 1)
// `fbt` = g_fb_timing read ONCE at kernel start: a mark is a clock read and a fire-and-forget store (re-reading the pointer from global memory
// at every mark put an L2 round trip, 500+ cycles, into the issuer's serial path and into the numbers)
#define FB_MARK(who, idx)                                                                \
    do {                                                                                 \
        if (fbt && (idx) < FT_MAX_MARKS) fbt[(who) * FT_MAX_MARKS + (idx)++] = clock64(); \
    } while (0)
#else
#define FT_MARK() \
    do {          \
    } while (0)
#define FB_MARK(who, idx) \
    do {                  \
    } while (0)
#define FB_SKIP(bit) 0
#define FB_CTA_MARK(i) \
    do {               \
    } while (0)
#endif

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
#define FS_AH1 16
#define FS_AH2 24
#define FS_CHUNKS_HEAD 32   // tile size when the pred-normals network gets no gradient (NeRF-VO: pred_normal_loss_mult = 0)
#define FS_P 32     // pred-normals input (32)
#define FS_AP1 36
#define FS_AP2 44
#define FS_AP3 52
#define FS_CHUNKS 60

struct FieldFwdP {
    int64_t n;
    int S;
    int want_pn, want_normals;
    int save_pn;                  // saved tiles carry the pred-normals activations too (FS_CHUNKS chunks instead of FS_CHUNKS_HEAD)
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
#define group_sync(g) \
    do {              \
        FT_MARK();    \
        group_sync_(g); \
        FT_MARK();    \
    } while (0)
__device__ __forceinline__ void group_sync_(int g) {
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

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
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
// sin(u) for the reference's fp32 argument u = fl(fl(2 pi x) 2^k) (+ fl(pi/2) for the cos block): two-term Cody-Waite reduction to [-pi, pi]
// and the SFU sine (abs. error < 1e-6 there) — the value is rounded to fp16 (5e-4) right afterwards
__device__ __forceinline__ float ft_sin_reduced(float u) {
    const float k = rintf(u * 0.15915494309189535f);
    const float ur = fmaf(-k, 1.7484555e-7f, fmaf(-k, 6.2831855f, u));  // 2 pi = 6.2831855 (fp32) - 1.7484555e-7
    return __sinf(ur);
}
__device__ __forceinline__ float ft_posenc(float xi, int k, bool cos_block) {
    const float u = __fmul_rn(__fmul_rn(6.283185307179586f, xi), (float)(1 << k));
    return ft_sin_reduced(cos_block ? __fadd_rn(u, 1.5707963267948966f) : u);
}

// ================================================================================================================================
// forward
// shared memory: [weight image FI_BYTES][ones tile 2 chunks][per group: FI 4 | H 8 | X 8 | P 4 chunks][mbarriers][tmem ptr]
// TMEM: the CTA owns all 512 columns (base 0); group g uses [128 g, 128 g + 128): accA = +[0,64) wide layers, accB = +[64,80) the narrow
// output layers (issued together with the next wide one), accN = +[96,128) the normals chain's input-gradient product
// ================================================================================================================================
#define FT_GROUP_CHUNKS 24
#define FT_ONES_OFF ((FI_BYTES + 1023) & ~1023)
#define FT_GROUPS_OFF (FT_ONES_OFF + 2 * CHUNK_B)
#define FT_MBAR_OFF(G) (FT_GROUPS_OFF + (G) * FT_GROUP_CHUNKS * CHUNK_B)

__device__ __forceinline__ void umma_commit_u32(uint32_t mbar_saddr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_saddr) : "memory");
}

// Shared-memory descriptor of the canonical no-swizzle layout at byte offset `off` (compile-time) from the CTA's shared-memory base, whose
// 16-byte index `base16` = (address & 0x3FFFF) >> 4 is the only run-time input: low word = base16 + constant (no carry into the LBO field:
// shared-memory addresses stay below 2^18), high word constant.  One uniform add per descriptor instead of a shift / mask / or chain.
__device__ __forceinline__ uint64_t umma_desc_rel(uint32_t base16, uint32_t off, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t lo = base16 + ((off >> 4) + ((lbo_bytes >> 4) << 16));
    const uint32_t hi = (sbo_bytes >> 4) | (1u << 14);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void issue_layer_rel(uint32_t tmem_d, uint32_t base16, uint32_t a_off, int ksteps, uint32_t w_off, int N, bool bias) {
    const uint32_t idesc = umma_idesc(TM, N, 0, 0);
#pragma unroll
    for (int k = 0; k < ksteps; ++k)
        umma_f16(tmem_d, umma_desc_rel(base16, a_off + k * 2 * CHUNK_B, CHUNK_B, 128), umma_desc_rel(base16, w_off + k * 2 * N * 16, N * 16, 128), idesc, k > 0);
    if (bias) umma_f16(tmem_d, umma_desc_rel(base16, FT_ONES_OFF, CHUNK_B, 128), umma_desc_rel(base16, w_off + ksteps * 2 * N * 16, N * 16, 128), idesc, 1);
}

// One step's MMAs of tile group GI.  Every operand is derived from the CTA's shared-memory base and compile-time constants (the group
// index included, TMEM base 0), so the descriptors live in uniform registers, one add each; with a run-time group index ptxas wraps every
// MMA in an elect / R2UR.BROADCAST loop behind a dependent shift / mask chain (~110 cycles per MMA instead of the pipe's 46).
template <int G, int GI>
__device__ __forceinline__ void ft_issue(int step, uint32_t uBase, int want_normals, int want_pn) {
    const uint32_t b16 = (uBase & 0x3FFFFu) >> 4;
    constexpr uint32_t oG = FT_GROUPS_OFF + GI * FT_GROUP_CHUNKS * CHUNK_B;
    constexpr uint32_t oFI = oG, oH = oG + 4 * CHUNK_B, oX = oG + 12 * CHUNK_B, oP = oG + 20 * CHUNK_B;
    constexpr uint32_t accA = GI * 128, accB = GI * 128 + 64, accN = GI * 128 + 96;
    switch (step) {
        case 0: issue_layer_rel(accA, b16, oFI, 2, FI_B0, 64, true); break;
        case 1:
            issue_layer_rel(accA, b16, oH, 4, FI_B1, 16, true);
            if (want_normals) issue_layer_rel(accN, b16, oX, 4, FI_NS, 32, false);
            break;
        case 2: issue_layer_rel(accA, b16, oX, 4, FI_H0, 64, false); break;
        case 3: issue_layer_rel(accA, b16, oH, 4, FI_H1, 64, true); break;
        case 4:
            issue_layer_rel(accB, b16, oH, 4, FI_H2, 16, true);
            if (want_pn) issue_layer_rel(accA, b16, oP, 2, FI_P0, 64, false);
            break;
        case 5: issue_layer_rel(accA, b16, oH, 4, FI_P1, 64, true); break;
        case 6: issue_layer_rel(accA, b16, oH, 4, FI_P2, 64, true); break;
        default: issue_layer_rel(accB, b16, oH, 4, FI_P3, 16, true); break;
    }
    umma_commit_u32(uBase + FT_MBAR_OFF(G) + 8 * (1 + GI));
}

template <int G>
__global__ void __launch_bounds__(G* FT_THREADS, 1) k_field_fwd(const __grid_constant__ FieldFwdP p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int g = threadIdx.x >> 8, tid = threadIdx.x & 255, warp = tid >> 5, hf = warp >> 2;
    const int r = ((warp & 3) << 5) | (tid & 31);  // row of the tile == TMEM lane
    unsigned char* sW = smem;
    unsigned char* sOnes = smem + FT_ONES_OFF;
    unsigned char* sG = smem + FT_GROUPS_OFF + g * FT_GROUP_CHUNKS * CHUNK_B;
    unsigned char *sFI = sG, *sH = sG + 4 * CHUNK_B, *sX = sG + 12 * CHUNK_B, *sP = sG + 20 * CHUNK_B;
    uint64_t* mbars = reinterpret_cast<uint64_t*>(smem + FT_MBAR_OFF(G));
    uint64_t* mbar_w = mbars;
    uint64_t* mbar_mma = mbars + 1 + g;
    uint64_t* mbar_in = mbars + 1 + G + g;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(mbars + 1 + 2 * G);
    const int64_t n_tiles = (p.n + TM - 1) >> 7;
    const int64_t stride = (int64_t)gridDim.x * G;
    const int64_t first = (int64_t)blockIdx.x * G + g;
    const uint32_t uBase = smem_u32(smem);

    // ones tile: chunk 0 = feature 0 is 1.0 for every row, chunk 1 = zeros
    for (int e = threadIdx.x; e < 2 * CHUNK_B / 16; e += G * FT_THREADS)
        reinterpret_cast<uint4*>(sOnes)[e] = e < TM ? make_uint4(0x00003C00u, 0, 0, 0) : make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
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
    if (*tmem_ptr != 0u) __trap();  // all 512 columns: the allocation can only start at column 0 (ft_issue relies on it)
    if (threadIdx.x == 0) {
        mbar_expect_tx(mbar_w, FI_BYTES);
        bulk_g2s(sW, p.wimg, FI_BYTES, mbar_w);
    }
    if (tid == 0 && first < n_tiles) {
        mbar_expect_tx(mbar_in, 4 * CHUNK_B);
        bulk_g2s(sFI, p.feat16 + first * (4 * CHUNK_B), 4 * CHUNK_B, mbar_in);
    }
    const uint32_t trow = (uint32_t)(g * 128) + ((uint32_t)((warp & 3) * 32) << 16);  // this group's columns, this warp's lanes
    mbar_wait(mbar_w, 0);
    uint32_t ph = 0;
    int it = 0;
#ifdef NVO_FT_TIMING
    int ft_mark = 0;
#endif
#define FT_ISSUE(step)                                                           \
    if (warp == 0) { /* warp-uniform; the group's first warp is converged here */ \
        tc_fence_after();                                                        \
        if (g == 0) {                                                            \
            if (elect_one()) ft_issue<G, 0>(step, uBase, p.want_normals, p.want_pn); \
        } else if (G > 1 && g == 1) {                                            \
            if (elect_one()) ft_issue<G, (G > 1 ? 1 : 0)>(step, uBase, p.want_normals, p.want_pn); \
        } else if (G > 2) {                                                      \
            if (elect_one()) ft_issue<G, (G > 2 ? 2 : 0)>(step, uBase, p.want_normals, p.want_pn); \
        }                                                                        \
        __syncwarp();                                                            \
    }
#define FT_COMMIT_WAIT()                   \
    FT_MARK();                             \
    mbar_wait(mbar_mma, ph);               \
    ph ^= 1;                               \
    tc_fence_after();                      \
    FT_MARK();

    for (int64_t tile = first; tile < n_tiles; tile += stride, ++it) {
        const int64_t t = tile * TM + r;
        const bool live = t < p.n;
        const int64_t tc = live ? t : p.n - 1;  // rows past the end read the last sample's inputs (finite) and write nothing
        uint4* sv = p.saved ? p.saved + tile * (int64_t)((p.save_pn ? FS_CHUNKS : FS_CHUNKS_HEAD) * TM) : nullptr;
        uint4* svp = p.save_pn ? sv : nullptr;  // pred-normals activations
        FT_MARK();
        mbar_wait(mbar_in, (uint32_t)(it & 1));
        // ---- S0: mlp_base layer 0 --------------------------------------------------------------------------------------------------
        FT_ISSUE(0);
        // per-ray inputs of the head's input tile: requested now, consumed two steps later (their latency hides behind S0 and S1)
        const int ray = (int)((uint32_t)tc / (uint32_t)p.S);
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, e0 = 0.f, selv = 0.f;
        uint4 x4, x5, x6, x7;
        if (hf == 0) {
            d0 = __ldg(p.dirs + 3 * ray), d1 = __ldg(p.dirs + 3 * ray + 1), d2 = __ldg(p.dirs + 3 * ray + 2);
            e0 = __ldg((p.cam ? p.emb + APP * __ldg(p.cam + ray) : p.emb));
            selv = __ldg(p.sel + tc);
        } else {
            // head input columns 32..63: appearance embedding 1..31, then the constant 1 that carries the first layer's bias
            const float* e = p.cam ? p.emb + APP * __ldg(p.cam + ray) : p.emb;
            float ev[32];
#pragma unroll
            for (int k = 0; k < APP; k += 4) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(e + k));
                ev[k] = q.x, ev[k + 1] = q.y, ev[k + 2] = q.z, ev[k + 3] = q.w;
            }
            x4 = pack8f(ev + 1), x5 = pack8f(ev + 9), x6 = pack8f(ev + 17);
            x7 = make_uint4(cvt_h2(ev[25], ev[26]), cvt_h2(ev[27], ev[28]), cvt_h2(ev[29], ev[30]), cvt_h2(ev[31], 1.f));
        }
        FT_COMMIT_WAIT();
        if (p.want_normals && hf == 1) {
            // the saved feature derivatives are read two steps later (normals): pull their 12 x 512 B per warp into L2 now
#pragma unroll
            for (int c = 0; c < 12; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.jac + (tile * 12 + c) * TM + r));
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
                    const uint32_t* w2 = reinterpret_cast<const uint32_t*>(&w);
                    uint32_t* dd = reinterpret_cast<uint32_t*>(&d);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dd[j] = w2[j] & __hgt2_mask(a2[j], zero);
                    *reinterpret_cast<uint4*>(sX + (4 * hf + q) * CHUNK_B + r * 16) = d;
                }
            }
        }
        group_sync(g);
        // the feature tile has been consumed (S0's MMA retired) and every thread of the group is past this tile's wait on mbar_in (the barrier
        // above) — only now may the barrier's next phase start: a thread still to make that wait would otherwise see the parity wrap around
        if (warp == 0 && tile + stride < n_tiles) {
            if (elect_one()) {
                mbar_expect_tx(mbar_in, 4 * CHUNK_B);
                bulk_g2s(sFI, p.feat16 + (tile + stride) * (4 * CHUNK_B), 4 * CHUNK_B, mbar_in);
            }
            __syncwarp();
        }
        // ---- S1: mlp_base layer 1 (16 outputs) + the normals chain's input-gradient product (32 feature columns, accN) ----------------------
        FT_ISSUE(1);
        FT_COMMIT_WAIT();
        float hv[16];
        tmem_ld16(trow, hv);  // raw density + geometry features (both halves need them)
        if (hf == 0) {
            // head input columns 0..31: SH16 | geo15 | appearance channel 0 (nerfacto_field.py:253-262)
            float c[16];
            ft_sh16(__fmul_rn(__fadd_rn(d0, 1.f), 0.5f), __fmul_rn(__fadd_rn(d1, 1.f), 0.5f), __fmul_rn(__fadd_rn(d2, 1.f), 0.5f), c);
            const uint4 x0 = pack8f(c), x1 = pack8f(c + 8), x2 = pack8f(hv + 1);
            const uint4 x3 = make_uint4(cvt_h2(hv[9], hv[10]), cvt_h2(hv[11], hv[12]), cvt_h2(hv[13], hv[14]), cvt_h2(hv[15], e0));
            // (the normals MMA read sX as its operand: it retired before the commit this epilogue waited for)
            *reinterpret_cast<uint4*>(sX + 0 * CHUNK_B + r * 16) = x0;
            *reinterpret_cast<uint4*>(sX + 1 * CHUNK_B + r * 16) = x1;
            *reinterpret_cast<uint4*>(sX + 2 * CHUNK_B + r * 16) = x2;
            *reinterpret_cast<uint4*>(sX + 3 * CHUNK_B + r * 16) = x3;
            if (sv) sv[(FS_X + 0) * TM + r] = x0, sv[(FS_X + 1) * TM + r] = x1, sv[(FS_X + 2) * TM + r] = x2, sv[(FS_X + 3) * TM + r] = x3;
        } else {
            *reinterpret_cast<uint4*>(sX + 4 * CHUNK_B + r * 16) = x4;
            *reinterpret_cast<uint4*>(sX + 5 * CHUNK_B + r * 16) = x5;
            *reinterpret_cast<uint4*>(sX + 6 * CHUNK_B + r * 16) = x6;
            *reinterpret_cast<uint4*>(sX + 7 * CHUNK_B + r * 16) = x7;
            if (sv) sv[(FS_X + 4) * TM + r] = x4, sv[(FS_X + 5) * TM + r] = x5, sv[(FS_X + 6) * TM + r] = x6, sv[(FS_X + 7) * TM + r] = x7;
        }
        group_sync(g);
        // ---- S2: mlp_head layer 0; behind its issue, off the chain: density, the pred-normals input tile, the density-gradient normals -----
        FT_ISSUE(2);
        if (hf == 0) {
            // density = trunc_exp(h0) * selector (nerfacto_field.py:216-221)
            if (live) {
                p.density[t] = __fmul_rn(expf(hv[0]), selv);
                if (p.h0) p.h0[t] = hv[0];
            }
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
                if (svp) svp[(FS_P + 0) * TM + r] = p0, svp[(FS_P + 1) * TM + r] = p1, svp[(FS_P + 2) * TM + r] = p2, svp[(FS_P + 3) * TM + r] = p3;
            }
        } else if (p.want_normals) {
            float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint4* jb = p.jac + ((tile * 4 + c) * 3) * TM + r;
                const uint4 J[3] = {__ldg(jb), __ldg(jb + TM), __ldg(jb + 2 * TM)};
                float v[8];
                tmem_ld8(trow + 96 + 8 * c, v);  // already x level scale (folded into the weights)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float g0 = v[2 * q], g1 = v[2 * q + 1];
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
        FT_COMMIT_WAIT();
        epi_hidden<true>(trow, sH, sv ? sv + FS_AH1 * TM : nullptr, hf, r);
        group_sync(g);
        // ---- S3: mlp_head layer 1 ---------------------------------------------------------------------------------------------------------
        FT_ISSUE(3);
        FT_COMMIT_WAIT();
        epi_hidden<true>(trow, sH, sv ? sv + FS_AH2 * TM : nullptr, hf, r);
        group_sync(g);
        // ---- S4 (+ S5): colour output layer, issued together with the pred-normals input layer ---------------------------------------------
        FT_ISSUE(4);
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
            epi_hidden<true>(trow, sH, svp ? svp + FS_AP1 * TM : nullptr, hf, r);
            group_sync(g);
            // ---- S6 / S7: mlp_pred_normals layers 1, 2 (the last one has no activation, mlp.py:143-179 with out_activation None) -------------
            FT_ISSUE(5);
            FT_COMMIT_WAIT();
            epi_hidden<true>(trow, sH, svp ? svp + FS_AP2 * TM : nullptr, hf, r);
            group_sync(g);
            FT_ISSUE(6);
            FT_COMMIT_WAIT();
            epi_hidden<false>(trow, sH, svp ? svp + FS_AP3 * TM : nullptr, hf, r);
            group_sync(g);
            // ---- S8: PredNormalsFieldHead: Linear(64, 3) + Tanh + normalize (field_heads.py:189-204) -------------------------------------------
            FT_ISSUE(7);
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
#undef FT_ISSUE
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(0u) : "memory");
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

__global__ void k_field_pack_bwd(const float* __restrict__ base, const float* __restrict__ head, unsigned char* __restrict__ img);

// ================================================================================================================================
extern "C" int64_t nvo_field_wimage_bytes(void) { return ((FI_BYTES + 127) & ~127) + 24576; }  // forward section + backward section (FB_OFF + FB_BYTES)
extern "C" int64_t nvo_field_saved_bytes(int64_t n, int32_t save_pn) { return ((n + TM - 1) / TM) * (int64_t)(save_pn ? FS_CHUNKS : FS_CHUNKS_HEAD) * CHUNK_B; }

extern "C" int nvo_field_pack_weights(void* stream, const nvo_grid_desc* grid, const float* base_params, const float* head_params, const float* pn_params,
                                      void* wimage) {
    NVO_CHECK(grid && grid->n_levels == 16, "field_pack_weights: the fused field kernels are specialised for the 16-level main grid");
    NVO_CHECK(base_params && head_params && wimage, "field_pack_weights: null pointer");
    GridScales sc;
    for (int i = 0; i < 16; ++i) sc.s[i] = grid->scalings[i];
    k_field_pack<<<16, 256, 0, (cudaStream_t)stream>>>(base_params, head_params, pn_params, sc, (unsigned char*)wimage);
    NVO_CUDA_LAUNCH_CHECK("field_pack_weights");
    k_field_pack_bwd<<<8, 256, 0, (cudaStream_t)stream>>>(base_params, head_params, (unsigned char*)wimage);
    NVO_CUDA_LAUNCH_CHECK("field_pack_weights(backward section)");
    return 0;
}

static int field_groups() {
    static const int g = nvo_env_int("NVO_FIELD_GROUPS", 3);
    return g >= 1 && g <= 3 ? g : 3;
}

extern "C" int nvo_field_forward(void* stream, int64_t B, int32_t S, const void* feat16, const void* jac, const float* positions, const float* directions,
                                 const int64_t* cam_idx, const float* embedding, const float* selector, const void* wimage, float* density, float* rgb,
                                 float* pred_normals, float* normals, float* h0, float* pn_raw, void* saved, int32_t save_pn) {
    NVO_CHECK(B >= 0 && S >= 1, "field_forward: bad shape B=%lld S=%d", (long long)B, S);
    if (B == 0) return 0;
    NVO_CHECK(feat16 && directions && embedding && selector && wimage && density && rgb, "field_forward: null pointer");
    NVO_CHECK(!pred_normals || positions, "field_forward: positions required for the predicted normals");
    NVO_CHECK(!normals || jac, "field_forward: the saved feature derivatives (nvo_grid_forward_jac) are required for the normals");
    FieldFwdP p;
    p.n = B * S, p.S = S, p.want_pn = pred_normals != nullptr, p.want_normals = normals != nullptr;
    p.feat16 = (const unsigned char*)feat16, p.jac = (const uint4*)jac, p.pos = positions, p.dirs = directions, p.cam = cam_idx, p.emb = embedding;
    p.sel = selector, p.wimg = (const unsigned char*)wimage, p.density = density, p.rgb = rgb, p.pn = pred_normals, p.normals = normals, p.h0 = h0;
    p.pn_raw = pn_raw, p.saved = (uint4*)saved, p.save_pn = (saved && save_pn && pred_normals) ? 1 : 0;
    const int G = field_groups();
    const size_t smem = FT_GROUPS_OFF + (size_t)G * FT_GROUP_CHUNKS * CHUNK_B + 8 * (1 + 2 * G) + 16;
    const int64_t tiles = (p.n + TM - 1) / TM;
    const unsigned int grid = (unsigned int)min((int64_t)nvo_sm_count(), (tiles + G - 1) / G);
    cudaError_t e;
    if (G == 1) {
        e = cudaFuncSetAttribute(k_field_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        NVO_CHECK(e == cudaSuccess, "field_forward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        k_field_fwd<1><<<grid, FT_THREADS, smem, (cudaStream_t)stream>>>(p);
    } else if (G == 2) {
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

// ================================================================================================================================
// backward: mlp_head (dgrad + wgrad), input assembly backward (geometry features, appearance-embedding gradient, trunc_exp'), mlp_base
// (dgrad + wgrad) -> d loss / d hash features (fp32 TMF tiles for the table scatter) in ONE persistent kernel.
//
//   * G tile groups of 256 epilogue threads + one MMA-issuing warp.  The weight gradients of all five layers stay resident in TMEM for the
//     whole kernel (dW^T[in][out] += A_ext^T dZ, K = the tile's 128 samples, the bias as the row of a constant ones feature) and are SHARED by
//     the groups: every tcgen05.mma of the CTA is issued by the same thread, so accumulation into the same columns is ordinary in-order
//     accumulation.  A group signals "operands ready" on an mbarrier (one arrival per warp), the issuer polls the groups round-robin, issues
//     the step's dgrad + wgrad MMAs and commits to the group's "done" mbarrier.
//   * per tile 5 steps: head L2 | head L1 | head L0 | base L1 | base L0.  Saved activations arrive just in time: two 18 KB slots per group
//     (tile + ones chunk), the issuer refills slot (q+1) & 1 with step q+1's tile by a bulk copy when step q's operands are ready (its last
//     readers, step q-1's MMAs and epilogue, are finished by then).
//   * gradients are scaled by powers of two before their fp16 conversion and unscaled in fp32 on the way out (tiny-cuda-nn's loss scale,
//     TCNN/include/tiny-cuda-nn/common.h:232), with TWO scales chosen per launch by k_field_bwd_absmax: s_h for mlp_head's chain from
//     max|d rgb sigmoid'|, s_b for mlp_base's chain from max(|d density trunc_exp'|, 4 max|d rgb sigmoid'|).  One scale for both would push
//     the colour path's gradients into fp16 subnormals whenever the density path's are much larger.  The largest scaled entry starts in
//     [128, 256): a chain would have to amplify it 256-fold to leave fp16's range, and the conversions saturate instead of producing inf.
//     (kind::f16 does not take fp16 activations with bf16 gradients: mixed operand formats raise an illegal-instruction fault on B200.)
// TMEM: [0,16) dW_head2^T | [16,32) dW_base1^T | [32,96) dW_head1^T | [96,160) dW_head0^T | [160,224) dW_base0^T | 224 + 64 g: group g's dgrad
// ================================================================================================================================
#define FB_OFF ((FI_BYTES + 127) & ~127)
#define FB_H2T 0                         // N 64 (inputs), K 16 (outputs)
#define FB_H1T (FB_H2T + 64 * 16 * 2)    // N 64, K 64
#define FB_H0T (FB_H1T + 64 * 64 * 2)    // N 64, K 64 (input 63 = the bias column: zero row)
#define FB_B1T (FB_H0T + 64 * 64 * 2)    // N 64, K 16
#define FB_B0T (FB_B1T + 64 * 16 * 2)    // N 32, K 64
#define FB_BYTES (FB_B0T + 32 * 64 * 2)

#define BW_SLOT_CHUNKS 9
#define BW_GROUP_CHUNKS (2 * BW_SLOT_CHUNKS + 8 + 2)
#define BW_GROUPS_OFF ((FB_BYTES + 1023) & ~1023)
#define DW_H2T 0
#define DW_B1T 16
#define DW_H1T 32
#define DW_H0T 96
#define DW_B0T 160
#define DW_ACC 224

struct FieldBwdP {
    int64_t n;
    int S;
    int saved_chunks;             // chunks per saved tile (FS_CHUNKS_HEAD or FS_CHUNKS)
    const unsigned char* feat16;  // TMH [tiles][4][128][8]
    const unsigned char* saved;   // forward's saved tiles
    const unsigned char* wimg;    // weight image (backward section at FB_OFF)
    const float* rgb;             // [n,3] forward output (sigmoid')
    const float* h0;              // [n]
    const float* sel;             // [n]
    const int64_t* cam;           // [B] or null
    const float* ddensity;        // [n] or null
    const float* drgb;            // [n,3]
    const float* dpn_in;          // fp32 TMF [tiles][27][128] (gradient w.r.t. the pred-normals input, from its own backward) or null
    const float* absmax;          // device float[2]: max |d rgb sigmoid'|, max |d density trunc_exp' selector|
    float* dfeat;                 // fp32 TMF [tiles][32][128]
    float* dbase;                 // flat fp32 gradient of mlp_base (torch layout), accumulated
    float* dhead;                 // flat fp32 gradient of mlp_head, accumulated
    float* demb;                  // [K,32] (cam != null) or [32], accumulated; nullable
    const float* dirs;            // [B,3] ray directions (with ddirs)
    float* ddirs;                 // [B,3] d loss / d ray direction through the SH encoding, accumulated (atomics); nullable
    int l2_prefetch;              // prefetch the next tile's saved activations into L2 (NVO_FIELD_BWD_PREFETCH, default 1)
};

__device__ __forceinline__ float ft_grad_scale(float mx) {
    if (!(mx > 0.f) || !isfinite(mx)) return 1.f;
    int e;
    frexpf(mx, &e);             // mx = f * 2^e, f in [0.5, 1)
    return ldexpf(1.f, 8 - e);  // mx * scale in [128, 256)
}
__device__ __forceinline__ uint32_t cvt_sat_h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint4 pack8sat(const float* v) {
    return make_uint4(cvt_sat_h2(v[0], v[1]), cvt_sat_h2(v[2], v[3]), cvt_sat_h2(v[4], v[5]), cvt_sat_h2(v[6], v[7]));
}

// out[0] = max |d rgb sigmoid'(rgb)|, out[1] = max |d density selector trunc_exp'(h0)| (non-negative floats order like their bit patterns)
__global__ void __launch_bounds__(256) k_field_bwd_absmax(int64_t n, const float* __restrict__ drgb, const float* __restrict__ rgb,
                                                          const float* __restrict__ ddensity, const float* __restrict__ sel,
                                                          const float* __restrict__ h0, float* __restrict__ out) {
    float m = 0.f, md = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float y = __ldg(rgb + 3 * i + j);
            m = fmaxf(m, fabsf(__ldg(drgb + 3 * i + j) * y * (1.f - y)));
        }
        if (ddensity) md = fmaxf(md, fabsf(__ldg(ddensity + i) * __ldg(sel + i) * expf(fminf(fmaxf(__ldg(h0 + i), -15.f), 15.f))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        md = fmaxf(md, __shfl_xor_sync(0xffffffffu, md, o));
    }
    // one pair of atomics per CTA (same-address atomics from every warp of the grid serialise in L2: 4736 x 2 of them cost more than the pass itself)
    __shared__ float sm[8], smd[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m, smd[threadIdx.x >> 5] = md;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; ++w) m = fmaxf(m, sm[w]), md = fmaxf(md, smd[w]);
        if (m > 0.f) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
        if (md > 0.f) atomicMax(reinterpret_cast<int*>(out) + 1, __float_as_int(md));
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(mbar)) : "memory"); }

// wgrad, transposed form: D[feature i (M = 128: 64 real + ones row + ignored rows)][output o (N)] += sum over the tile's 128 samples of
// A_ext[s][i] * dZ[s][o]; both operands MN-major (the activation / dZ tiles as they lie in shared memory), 8 K-steps of 16 samples
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_d, uint32_t act_base, uint32_t dz_base, int N, uint32_t accumulate) {
    const uint32_t idesc = umma_idesc(TM, N, 1, 1);
    for (int k = 0; k < TM / 16; ++k)
        umma_f16(tmem_d, umma_desc(act_base + k * 256, 128, CHUNK_B), umma_desc(dz_base + k * 256, 128, CHUNK_B), idesc, k > 0 ? 1u : accumulate);
}

// dgrad epilogue of a hidden layer: 32 accumulator columns x ReLU'(saved activation) -> fp16 dZ tile
__device__ __forceinline__ void epi_mask(uint32_t tacc, const unsigned char* __restrict__ sAct, unsigned char* __restrict__ sOut, int hf, int r) {
    float v[32];
    tmem_ld32(tacc + 32 * hf, v);
    const __half2 zero = __float2half2_rn(0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 a = *reinterpret_cast<const uint4*>(sAct + (4 * hf + q) * CHUNK_B + r * 16);
        uint4 u = pack8sat(v + 8 * q);
        const __half2* a2 = reinterpret_cast<const __half2*>(&a);
        uint32_t* u2 = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) u2[j] &= __hgt2_mask(a2[j], zero);  // 0xFFFF per half where the saved fp16 activation is positive
        *reinterpret_cast<uint4*>(sOut + (4 * hf + q) * CHUNK_B + r * 16) = u;
    }
}

// sum over the lanes of a warp of e[k] for every k: afterwards e[0] of lane k holds channel k's total (31 shuffles)
__device__ __forceinline__ void warp_transpose_reduce32(float* e, int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float send = up ? e[i] : e[i + o];
            const float keep = up ? e[i + o] : e[i];
            e[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
}

// The backward's MMA issue with every operand address derived from the CTA's shared-memory base + compile-time constants (group, slot and
// step are template / switch constants, TMEM base 0): descriptors stay in uniform registers, one add each.  With run-time addresses ptxas
// wraps every tcgen05.mma of the single issuing thread in an elect / R2UR.BROADCAST loop behind a shift / mask chain (~110 cycles per MMA
// against the pipe's 32-46), and that one thread issues the 54 MMAs per tile of all groups: it was the CTA's serial bottleneck.
__device__ __forceinline__ void issue_wgrad_rel(uint32_t tmem_d, uint32_t base16, uint32_t act_off, uint32_t dz_off, int N, uint32_t accumulate) {
    const uint32_t idesc = umma_idesc(TM, N, 1, 1);
#pragma unroll
    for (int k = 0; k < TM / 16; ++k)
        umma_f16(tmem_d, umma_desc_rel(base16, act_off + k * 256, 128, CHUNK_B), umma_desc_rel(base16, dz_off + k * 256, 128, CHUNK_B), idesc, k > 0 ? 1u : accumulate);
}
#define BW_MBAR_OFF(G) (BW_GROUPS_OFF + ((G) * BW_GROUP_CHUNKS + 2) * CHUNK_B)
// One step's MMAs + commit with the step TS a compile-time constant as well (the static issue schedule of k_field_bwd): no switch, the
// elected lane's instruction stream is the uniform-register descriptor adds and the tcgen05.mma themselves.
template <int G, int GI, int SL, int TS>
__device__ __forceinline__ void fb_issue_s(uint32_t uBase, uint32_t started, uint32_t signal_go) {
    const uint32_t b16 = (uBase & 0x3FFFFu) >> 4;
    constexpr uint32_t oG = BW_GROUPS_OFF + GI * BW_GROUP_CHUNKS * CHUNK_B, oSlot = oG + SL * BW_SLOT_CHUNKS * CHUNK_B;
    constexpr uint32_t oG64 = oG + 2 * BW_SLOT_CHUNKS * CHUNK_B, oG16 = oG64 + 8 * CHUNK_B;
    constexpr uint32_t acc = DW_ACC + 64 * GI;
    if (elect_one()) {
        if (signal_go) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(uBase + BW_MBAR_OFF(G) + 8 * (1 + 4 * G + GI)) : "memory");
        if constexpr (TS == 0) {  // head layer 2: dA2 = dz3 W2 ; dW2^T += AH2_ext^T dz3
            if (!FB_SKIP(1)) issue_layer_rel(acc, b16, oG16, 1, FB_H2T, 64, false);
            if (!FB_SKIP(0)) issue_wgrad_rel(DW_H2T, b16, oSlot, oG16, 16, started);
        } else if constexpr (TS == 1) {  // head layer 1
            if (!FB_SKIP(1)) issue_layer_rel(acc, b16, oG64, 4, FB_H1T, 64, false);
            if (!FB_SKIP(0)) issue_wgrad_rel(DW_H1T, b16, oSlot, oG64, 64, started);
        } else if constexpr (TS == 2) {  // head layer 0: dX ; dW0^T += X^T dZ1 (the bias is X's constant column 63)
            if (!FB_SKIP(1)) issue_layer_rel(acc, b16, oG64, 4, FB_H0T, 64, false);
            if (!FB_SKIP(0)) issue_wgrad_rel(DW_H0T, b16, oSlot, oG64, 64, started);
        } else if constexpr (TS == 3) {  // base layer 1
            if (!FB_SKIP(1)) issue_layer_rel(acc, b16, oG16, 1, FB_B1T, 64, false);
            if (!FB_SKIP(0)) issue_wgrad_rel(DW_B1T, b16, oSlot, oG16, 16, started);
        } else {  // base layer 0: d features (32 columns); the feature tile sits in the slot's upper half
            if (!FB_SKIP(1)) issue_layer_rel(acc, b16, oG64, 4, FB_B0T, 32, false);
            if (!FB_SKIP(0)) issue_wgrad_rel(DW_B0T, b16, oSlot + 4 * CHUNK_B, oG64, 64, started);
        }
        umma_commit_u32(uBase + BW_MBAR_OFF(G) + 8 * (1 + G + GI));
    }
    __syncwarp();
}

template <int G>
__global__ void __launch_bounds__(G* FT_THREADS + 64, 1) k_field_bwd(const __grid_constant__ FieldBwdP p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int NEPI = G * FT_THREADS;
    const bool is_issuer_warp = threadIdx.x >= NEPI && threadIdx.x < NEPI + 32, is_loader_warp = threadIdx.x >= NEPI + 32;
    const bool is_group_thread = threadIdx.x < NEPI;
    const int g = is_group_thread ? (int)(threadIdx.x >> 8) : 0;
    const int tid = threadIdx.x & 255, warp = tid >> 5, hf = warp >> 2, lane = threadIdx.x & 31;
    const int r = ((warp & 3) << 5) | lane;
    unsigned char* sW = smem;
    const uint32_t uBase = smem_u32(smem);  // defined in converged code: the issuer's descriptors derive from it in uniform registers
    uint64_t* mbars = reinterpret_cast<uint64_t*>(smem + BW_MBAR_OFF(G));
    uint64_t* mbar_w = mbars;                 // weights landed
    uint64_t* mbar_ready = mbars + 1;         // [G] operands of the group's next step are in shared memory (8 warp arrivals)
    uint64_t* mbar_done = mbars + 1 + G;      // [G] the step's MMAs have retired (tcgen05.commit)
    uint64_t* mbar_load = mbars + 1 + 2 * G;  // [G][2] the slot's bulk copy has landed
    uint64_t* mbar_go = mbars + 1 + 4 * G;    // [G] issuer -> loader: step q's operands are ready, the other slot may be refilled
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(mbars + 1 + 5 * G);
    const int64_t n_tiles = (p.n + TM - 1) >> 7;
    const int64_t stride = (int64_t)gridDim.x * G;
    const size_t saved_tile_bytes = (size_t)p.saved_chunks * CHUNK_B;
    FB_CTA_MARK(0);

    // zero the activation slots / dZ buffers once (padded chunks feed ignored accumulator rows, but must not hold NaN patterns that
    // would reach real rows through the K dimension), then the constant ones chunk behind every slot
    for (int e = threadIdx.x; e < (G * BW_GROUP_CHUNKS + 2) * CHUNK_B / 16; e += NEPI + 64)
        reinterpret_cast<uint4*>(smem + BW_GROUPS_OFF)[e] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (is_group_thread && tid < TM) {
        unsigned char* sG = smem + BW_GROUPS_OFF + g * BW_GROUP_CHUNKS * CHUNK_B;
        *reinterpret_cast<uint4*>(sG + 8 * CHUNK_B + tid * 16) = make_uint4(0x00003C00u, 0, 0, 0);
        *reinterpret_cast<uint4*>(sG + (BW_SLOT_CHUNKS + 8) * CHUNK_B + tid * 16) = make_uint4(0x00003C00u, 0, 0, 0);
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        mbar_init(mbar_w, 1);
        for (int i = 0; i < G; ++i) {
            mbar_init(mbar_ready + i, 8);
            mbar_init(mbar_done + i, 1);
            mbar_init(mbar_load + 2 * i, 1);
            mbar_init(mbar_load + 2 * i + 1, 1);
            mbar_init(mbar_go + i, 1);
        }
        fence_mbar_init();
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *tmem_ptr;
    if (tmem0 != 0u) __trap();  // all 512 columns: the allocation can only start at column 0 (fb_issue relies on it)
    FB_CTA_MARK(1);
#ifdef NVO_FT_TIMING
    unsigned long long* const fbt = blockIdx.x == 0 ? g_fb_timing : nullptr;
#endif
    const float mx_h = __ldg(p.absmax), mx_d = __ldg(p.absmax + 1);
    const float s_h = ft_grad_scale(mx_h), s_b = ft_grad_scale(fmaxf(mx_d, 4.f * mx_h));
    const float inv_s_h = 1.f / s_h, inv_s_b = 1.f / s_b, s_bh = s_b * inv_s_h;

    if (is_issuer_warp || is_loader_warp) {
        // ===================================== MMA issuer warp and loader warp =====================================
        // Both walk the SAME static schedule: the groups are served round-robin in a fixed order, one step each — two tiles (10 steps) per loop
        // iteration, so the step, the slot and the barrier parities of every block below are compile-time constants.
        //   issuer, per block: wait `ready` (operands written) and `load` (the slot's copy landed) -> arrive on `go` -> the step's MMAs -> commit.
        //   loader, per block: wait `go` -> request the bulk copy of step q + 1 into the other slot (its readers, step q - 1's MMAs and epilogue,
        //     are finished once step q's operands are ready) and the L2 prefetch of a piece of the group's next tile.
        // History (tools/field_timing.cu, profiles/r02_field_bwd_phase_timing*.log): a polling issuer (any ready group, run-time step) spent
        // 150-200 scalar instructions = 900-1300 cycles per step next to 24 epilogue warps, against 550-720 for the step's MMAs, and that one
        // warp serialises the groups; with the copy requests (64-bit address arithmetic, ~250 cycles) in a second warp the issuer's block is
        // the MMAs themselves.  `go` phases cannot run ahead of the loader: the issuer arrives on go(q + 1) only after load(q + 1) has landed,
        // which the loader requests after it has seen go(q) (q = 0 has no `go`: steps 0 and 1 are preloaded; the last step requests nothing).
        int64_t tiles_g[G];
        int64_t tiles_max = 0;
#pragma unroll
        for (int gg = 0; gg < G; ++gg) {
            const int64_t first = (int64_t)blockIdx.x * G + gg;
            tiles_g[gg] = first < n_tiles ? (n_tiles - first + stride - 1) / stride : 0;
            tiles_max = tiles_g[gg] > tiles_max ? tiles_g[gg] : tiles_max;
        }
#ifdef NVO_FT_TIMING
        int fb_mi = 0;
#endif
        if (is_loader_warp) {
            const bool l2_prefetch = p.l2_prefetch != 0;
            auto load_s = [&](int gg, int ts, int sl, int64_t tile_local) {  // the elected lane only; gg, ts, sl are constants at the call sites
                const int64_t tile = (int64_t)blockIdx.x * G + gg + tile_local * stride;
                unsigned char* slot = smem + BW_GROUPS_OFF + gg * BW_GROUP_CHUNKS * CHUNK_B + sl * BW_SLOT_CHUNKS * CHUNK_B;
                uint64_t* mb = mbar_load + 2 * gg + sl;
                const unsigned char* sv = p.saved + tile * saved_tile_bytes;
                if (ts == 4) {  // base layer 0's input: the hash features (4 chunks), placed so that the ones chunk follows them
                    mbar_expect_tx(mb, 4 * CHUNK_B);
                    bulk_g2s(slot + 4 * CHUNK_B, p.feat16 + tile * (4 * CHUNK_B), 4 * CHUNK_B, mb);
                } else {
                    const int ch = ts == 0 ? FS_AH2 : ts == 1 ? FS_AH1 : ts == 2 ? FS_X : FS_H1;
                    mbar_expect_tx(mb, 8 * CHUNK_B);
                    bulk_g2s(slot, sv + ch * CHUNK_B, 8 * CHUNK_B, mb);
                }
                // the group's NEXT tile starts towards L2 in four pieces, each behind a slot copy (the TMA engine serves its requests in order:
                // one 72 KB prefetch ahead of a copy delayed that copy past the step that needed it)
                if (tile + stride < n_tiles && l2_prefetch) {
                    const unsigned char* nx = p.saved + (tile + stride) * saved_tile_bytes;
                    if (ts == 1) bulk_prefetch_l2(nx + FS_AH2 * CHUNK_B, 8 * CHUNK_B);
                    if (ts == 2) bulk_prefetch_l2(nx + FS_AH1 * CHUNK_B, 8 * CHUNK_B);
                    if (ts == 3) bulk_prefetch_l2(nx + FS_X * CHUNK_B, 8 * CHUNK_B);
                    if (ts == 4) {
                        bulk_prefetch_l2(nx + FS_H1 * CHUNK_B, 8 * CHUNK_B);
                        bulk_prefetch_l2(p.feat16 + (tile + stride) * (4 * CHUNK_B), 4 * CHUNK_B);
                    }
                }
            };
            if (elect_one()) {
                mbar_expect_tx(mbar_w, FB_BYTES);
                bulk_g2s(sW, p.wimg + FB_OFF, FB_BYTES, mbar_w);
#pragma unroll
                for (int gg = 0; gg < G; ++gg)
                    if (tiles_g[gg] > 0) {
                        load_s(gg, 0, 0, 0);
                        load_s(gg, 1, 1, 0);
                    }
            }
            __syncwarp();
            for (int64_t it = 0; it < tiles_max; it += 2) {
#define FB_LBLOCK(GI, U)                                                                                                   \
    if (GI < G && it + ((U) >= 5) < tiles_g[GI < G ? GI : 0] && (it > 0 || (U) > 0)) {                                      \
        constexpr int gi = GI < G ? GI : 0, sl = (U) & 1, ts = (U) % 5;                                                    \
        mbar_wait(mbar_go + gi, (uint32_t)(((U) + 1) & 1));                                                                \
        const int64_t tl = it + ((U) >= 5);                /* this step's tile of the group */                             \
        if (ts < 4 || tl + 1 < tiles_g[gi]) {              /* step q + 1 exists */                                         \
            if (elect_one()) load_s(gi, (ts + 1) % 5, sl ^ 1, tl + (ts == 4));                                            \
            __syncwarp();                                                                                                 \
        }                                                                                                                 \
    }
#define FB_LROUND(U) FB_LBLOCK(0, U) FB_LBLOCK(1, U) FB_LBLOCK(2, U)
                FB_LROUND(0) FB_LROUND(1) FB_LROUND(2) FB_LROUND(3) FB_LROUND(4) FB_LROUND(5) FB_LROUND(6) FB_LROUND(7) FB_LROUND(8) FB_LROUND(9)
#undef FB_LROUND
#undef FB_LBLOCK
            }
        } else {
            mbar_wait(mbar_w, 0);
            uint32_t lbits = 0;  // bit 2 gg + sl: phase parity of the slot's load barrier
            for (int64_t it = 0; it < tiles_max; it += 2) {
#define FB_BLOCK(GI, U)                                                                                                   \
    if (GI < G && it + ((U) >= 5) < tiles_g[GI < G ? GI : 0]) {                                                            \
        constexpr int gi = GI < G ? GI : 0, sl = (U) & 1, ts = (U) % 5;                                                    \
        if (gi == 0 && lane == 0) FB_MARK(1, fb_mi);                                                                      \
        mbar_wait(mbar_ready + gi, (uint32_t)sl);                                                                         \
        if (gi == 0 && lane == 0) FB_MARK(1, fb_mi);                                                                      \
        mbar_wait(mbar_load + 2 * gi + sl, (lbits >> (2 * gi + sl)) & 1u);                                                \
        lbits ^= 1u << (2 * gi + sl);                                                                                     \
        if (gi == 0 && lane == 0) FB_MARK(1, fb_mi);                                                                      \
        tc_fence_after();                                                                                                 \
        fb_issue_s<G, gi, sl, ts>(uBase, (it > 0 || (U) >= 5 || gi > 0) ? 1u : 0u, (it > 0 || (U) > 0) ? 1u : 0u);         \
        if (gi == 0 && lane == 0) FB_MARK(1, fb_mi);                                                                      \
    }
#define FB_ROUND(U) FB_BLOCK(0, U) FB_BLOCK(1, U) FB_BLOCK(2, U)
                FB_ROUND(0) FB_ROUND(1) FB_ROUND(2) FB_ROUND(3) FB_ROUND(4) FB_ROUND(5) FB_ROUND(6) FB_ROUND(7) FB_ROUND(8) FB_ROUND(9)
#undef FB_ROUND
#undef FB_BLOCK
            }
        }
        __syncwarp();
    } else {
        // ==================================================== tile groups ====================================================
        unsigned char* sG = smem + BW_GROUPS_OFF + g * BW_GROUP_CHUNKS * CHUNK_B;
        unsigned char* sG64 = sG + 2 * BW_SLOT_CHUNKS * CHUNK_B;
        unsigned char* sG16 = sG64 + 8 * CHUNK_B;
        const uint32_t tacc = tmem0 + DW_ACC + 64 * g + ((uint32_t)((warp & 3) * 32) << 16);
        uint64_t* ready = mbar_ready + g;
        uint64_t* done = mbar_done + g;
        uint32_t dph = 0;
        int q = 0;
#ifdef NVO_FT_TIMING
        int fb_mg = 0;
        const bool fb_me = g == 0 && tid == 0;
#endif
        auto arrive_ready = [&]() {
#ifdef NVO_FT_TIMING
            if (fb_me) FB_MARK(0, fb_mg);  // epilogue of the previous step done
#endif
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ready);
#ifdef NVO_FT_TIMING
            if (fb_me) FB_MARK(0, fb_mg);  // arrived
#endif
        };
        auto wait_done = [&]() {
            mbar_wait(done, dph);
            dph ^= 1;
            tc_fence_after();
#ifdef NVO_FT_TIMING
            if (fb_me) FB_MARK(0, fb_mg);  // MMAs retired
#endif
            mbar_wait(mbar_load + 2 * g + (q & 1), (uint32_t)((q >> 1) & 1));  // already complete (the issuer waited for it): orders the reads below
        };
        // dz of the colour output: drgb * sigmoid'(rgb) * scale -> G16 chunk 0 (outputs 0..2), chunk 1 = 0.  Two halves: the global loads are
        // requested a step early (their latency hides behind the wait for the MMAs), the shared-memory write follows once G16's readers retired.
        auto load_dz3 = [&](int64_t tile, float* dz) {
            dz[0] = dz[1] = dz[2] = 0.f;
            if (hf != 0) return;
            const int64_t t = tile * TM + r;
            if (tile < n_tiles && t < p.n) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float y = __ldg(p.rgb + 3 * t + j);
                    dz[j] = __ldg(p.drgb + 3 * t + j) * y * (1.f - y) * s_h;
                }
            }
        };
        auto store_dz3 = [&](const float* dz3) {
            if (hf != 0) return;
            const float dz[8] = {dz3[0], dz3[1], dz3[2], 0.f, 0.f, 0.f, 0.f, 0.f};
            *reinterpret_cast<uint4*>(sG16 + r * 16) = pack8f(dz);
            *reinterpret_cast<uint4*>(sG16 + CHUNK_B + r * 16) = make_uint4(0, 0, 0, 0);
        };
        const int64_t first = (int64_t)blockIdx.x * G + g;
        {
            float dz0[3];
            load_dz3(first, dz0);
            store_dz3(dz0);
        }
        for (int64_t tile = first; tile < n_tiles; tile += stride) {
            const int64_t t = tile * TM + r;
            const bool live = t < p.n;
            const int64_t tc = live ? t : p.n - 1;
            // ---- step 0: head layer 2 -> dZ2 = dA2 * ReLU'(AH2) -------------------------------------------------------------------------
            arrive_ready();
            wait_done();
            epi_mask(tacc, sG + (q & 1) * BW_SLOT_CHUNKS * CHUNK_B, sG64, hf, r);
            ++q;
            // ---- step 1: head layer 1 -> dZ1 (in place: the step's MMAs, the buffer's readers, have retired) ---------------------------------
            arrive_ready();
            wait_done();
            epi_mask(tacc, sG + (q & 1) * BW_SLOT_CHUNKS * CHUNK_B, sG64, hf, r);
            ++q;
            // ---- step 2: head layer 0 -> dX: geometry features -> dh, appearance embedding gradient, trunc_exp' --------------------------------
            arrive_ready();
            // d loss / d raw density (global loads: requested before the wait for the MMAs)
            float dh0 = 0.f;
            if (hf == 0 && live && p.ddensity)
                dh0 = __ldg(p.ddensity + t) * __ldg(p.sel + t) * expf(fminf(fmaxf(__ldg(p.h0 + t), -15.f), 15.f)) * s_b;  // activations.py:38-41
            wait_done();
            {
                float v[32];
                tmem_ld32(tacc + 32 * hf, v);
                const int64_t ray = tc / p.S;
                const int ray0 = __shfl_sync(0xffffffffu, (int)ray, 0);
                const bool second = (int)ray != ray0;                       // the warp's 32 rows span at most two rays (S >= 32 enforced)
                const bool split = __any_sync(0xffffffffu, second);
                const int ray1 = __shfl_sync(0xffffffffu, (int)ray, 31);
                if (hf == 0) {
                    float dh[16];
                    dh[0] = dh0;
#pragma unroll
                    for (int k = 0; k < GEO; ++k) dh[1 + k] = v[16 + k] * s_bh;  // head-chain scale -> base-chain scale
                    if (p.dpn_in && live) {
#pragma unroll
                        for (int k = 0; k < GEO; ++k) dh[1 + k] += s_b * __ldg(p.dpn_in + ((tile * 27 + 12 + k) << 7) + r);
                    }
                    *reinterpret_cast<uint4*>(sG16 + r * 16) = pack8sat(dh);
                    *reinterpret_cast<uint4*>(sG16 + CHUNK_B + r * 16) = pack8sat(dh + 8);
                }
                // step 3's operands (dh in G16) are complete and this step's accumulator is in registers: hand over to the issuer NOW, the per-ray
                // reductions below (direction / appearance-embedding gradients: shuffles and global atomics) run while base layer 1's MMAs do
                ++q;
                arrive_ready();
                if (hf == 0) {
                    if (p.ddirs) {
                        // head input columns 0..15 = SH16((d + 1) / 2): d loss / d d = 0.5 J_SH^T dSH (NS/utils/math.py:45-78), summed per ray
                        const float x = __fmul_rn(__fadd_rn(__ldg(p.dirs + 3 * ray), 1.f), 0.5f), y = __fmul_rn(__fadd_rn(__ldg(p.dirs + 3 * ray + 1), 1.f), 0.5f),
                                    z = __fmul_rn(__fadd_rn(__ldg(p.dirs + 3 * ray + 2), 1.f), 0.5f);
                        const float xx = x * x, yy = y * y, zz = z * z;
                        const float a1 = 0.4886025119029199f, b = 1.0925484305920792f, c6 = 0.9461746957575601f, e8 = 0.5462742152960396f,
                                    f9 = 0.5900435899266435f, g10 = 2.890611442640554f, h11 = 0.4570457994644658f, i12 = 0.3731763325901154f,
                                    j14 = 1.445305721320277f;
                        const float* G = v;  // G[k] = d loss / d SH_k (x head-chain scale)
                        float gd[3];
                        gd[0] = a1 * G[3] + b * y * G[4] + b * z * G[7] + 2.f * e8 * x * G[8] + 6.f * f9 * x * y * G[9] + g10 * y * z * G[10] +
                                h11 * (5.f * zz - 1.f) * G[13] + 2.f * j14 * z * x * G[14] + f9 * (3.f * xx - 3.f * yy) * G[15];
                        gd[1] = a1 * G[1] + b * x * G[4] + b * z * G[5] - 2.f * e8 * y * G[8] + f9 * (3.f * xx - 3.f * yy) * G[9] + g10 * x * z * G[10] +
                                h11 * (5.f * zz - 1.f) * G[11] - 2.f * j14 * z * y * G[14] - 6.f * f9 * x * y * G[15];
                        gd[2] = a1 * G[2] + b * y * G[5] + 2.f * c6 * z * G[6] + b * x * G[7] + g10 * x * y * G[10] + 10.f * h11 * y * z * G[11] +
                                i12 * (15.f * zz - 3.f) * G[12] + 10.f * h11 * x * z * G[13] + j14 * (xx - yy) * G[14];
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const float val = live ? gd[a] * (0.5f * inv_s_h) : 0.f;
                            const float s0 = nvo_warp_sum(second ? 0.f : val);
                            const float s1 = split ? nvo_warp_sum(second ? val : 0.f) : 0.f;
                            if (lane == 0) {
                                atomicAdd(p.ddirs + 3 * ray0 + a, s0);
                                if (split) atomicAdd(p.ddirs + 3 * ray1 + a, s1);
                            }
                        }
                    }
                    if (p.demb) {  // appearance channel 0 (head input column 31)
                        float a = second ? 0.f : v[31], b = second ? v[31] : 0.f;
                        a = nvo_warp_sum(a);
                        if (split) b = nvo_warp_sum(b);
                        if (lane == 0) {
                            atomicAdd(p.demb + (p.cam ? APP * __ldg(p.cam + ray0) : 0), a * inv_s_h);
                            if (split) atomicAdd(p.demb + (p.cam ? APP * __ldg(p.cam + ray1) : 0), b * inv_s_h);
                        }
                    }
                } else if (p.demb) {
                    // appearance channels 1..31 = columns 32..62 (column 63 carries the bias): per-ray sums by a transpose-reduce, one
                    // atomic per (ray, channel) instead of one per (sample, channel)
                    float e[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) e[k] = second ? 0.f : v[k];
                    warp_transpose_reduce32(e, lane);
                    if (lane < 31) atomicAdd(p.demb + (p.cam ? APP * __ldg(p.cam + ray0) : 0) + 1 + lane, e[0] * inv_s_h);
                    if (split) {
#pragma unroll
                        for (int k = 0; k < 32; ++k) e[k] = second ? v[k] : 0.f;
                        warp_transpose_reduce32(e, lane);
                        if (lane < 31) atomicAdd(p.demb + (p.cam ? APP * __ldg(p.cam + ray1) : 0) + 1 + lane, e[0] * inv_s_h);
                    }
                }
            }
            // ---- step 3: base layer 1 -> dZ0 = dA * ReLU'(H1) (arrival: above) ---------------------------------------------------------------
            wait_done();
            epi_mask(tacc, sG + (q & 1) * BW_SLOT_CHUNKS * CHUNK_B, sG64, hf, r);
            ++q;
            // ---- step 4: base layer 0 -> d features (fp32 TMF [tile][32][128]: a warp stores 128 contiguous bytes per column) ------------------
            arrive_ready();
            float dz_next[3];
            load_dz3(tile + stride, dz_next);  // the next tile's colour gradient: in flight during the wait
            wait_done();
            store_dz3(dz_next);  // G16's readers (step 3's MMAs) have retired; written first, so the next tile's arrival only waits for the TMEM read
            {
                float v[16];
                tmem_ld16(tacc + 16 * hf, v);
                float* dst = p.dfeat + ((tile * 32 + 16 * hf) << 7) + r;
#pragma unroll
                for (int j = 0; j < 16; ++j) dst[j << 7] = v[j] * inv_s_b;
            }
            ++q;
        }
    }
    // ---- flush the weight gradients: lane = input feature (64 = the ones feature = bias), column = output -----------------------------------
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    FB_CTA_MARK(2);
    if (is_group_thread && (int64_t)blockIdx.x * G < n_tiles) {
        const uint32_t tlane = tmem0 + ((uint32_t)((warp & 3) * 32) << 16);
        // (matrix, TMEM column, outputs, inputs, bias feature row, destination weight / bias offsets, destination buffer)
        struct M {
            int col, N, K, brow, w_off, b_off, head;
        };
        constexpr M mats[5] = {
            {DW_H1T, 64, 64, 64, 64 * 63 + 64, 64 * 63 + 64 + 64 * 64, 1},
            {DW_H0T, 64, 63, 63, 0, 64 * 63, 1},
            {DW_B0T, 64, 32, 32, 0, 64 * 32, 0},
            {DW_H2T, 3, 64, 64, 64 * 63 + 64 + 64 * 64 + 64, 64 * 63 + 64 + 64 * 64 + 64 + 3 * 64, 1},
            {DW_B1T, 16, 64, 64, 64 * 32 + 64, 64 * 32 + 64 + 16 * 64, 0},
        };
#pragma unroll
        for (int m = 0; m < 5; ++m) {  // unrolled: the table above folds into immediates (as a run-time array it lived in local memory)
            if (m % G != g) continue;
            constexpr M zero = {0, 0, 0, 0, 0, 0, 0};
            const M y = m < 5 ? mats[m] : zero;
            float* dst = y.head ? p.dhead : p.dbase;
            const float inv = y.head ? inv_s_h : inv_s_b;
            if ((warp & 3) * 32 > y.brow) continue;  // this warp's lanes hold no real row
            const int ncol = y.N < 16 ? 16 : y.N;
            for (int c0 = hf * 16; c0 < ncol; c0 += 32) {
                float v[16];
                tmem_ld16(tlane + y.col + c0, v);
                if (r < y.K) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < y.N) atomicAdd(dst + y.w_off + (c0 + j) * y.K + r, v[j] * inv);
                } else if (r == y.brow) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < y.N) atomicAdd(dst + y.b_off + c0 + j, v[j] * inv);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    FB_CTA_MARK(3);
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem0) : "memory");
}

// backward section of the weight image: transposed tiles (N = inputs, K = outputs) for the dgrad products
__global__ void __launch_bounds__(256) k_field_pack_bwd(const float* __restrict__ base, const float* __restrict__ head, unsigned char* __restrict__ img) {
    struct T {
        int head, w_off, N_in, K_in, n_out, kpad, off;  // W[o][i]: o < n_out outputs, i < K_in inputs; tile rows N_in (inputs), K = kpad outputs
    };
    const T L[5] = {
        {1, 64 * 63 + 64 + 64 * 64 + 64, 64, 64, 3, 16, FB_H2T},
        {1, 64 * 63 + 64, 64, 64, 64, 64, FB_H1T},
        {1, 0, 64, 63, 64, 64, FB_H0T},
        {0, 64 * 32 + 64, 64, 64, 16, 16, FB_B1T},
        {0, 0, 32, 32, 64, 64, FB_B0T},
    };
    const int t = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int l = 0; l < 5; ++l) {
        const T& y = L[l];
        const float* src = y.head ? head : base;
        __half* iw = reinterpret_cast<__half*>(img + FB_OFF + y.off);
        for (int e = t; e < y.N_in * y.kpad; e += stride) {
            const int i = e / y.kpad, o = e - i * y.kpad;
            const float w = (i < y.K_in && o < y.n_out) ? __ldg(src + y.w_off + o * y.K_in + i) : 0.f;
            iw[((o >> 3) * y.N_in + i) * 8 + (o & 7)] = __float2half_rn(w);
        }
    }
}

extern "C" int nvo_field_backward(void* stream, int64_t B, int32_t S, const void* feat16, const void* saved, int32_t save_pn, const void* wimage,
                                  const float* rgb, const float* h0, const float* selector, const int64_t* cam_idx, const float* ddensity,
                                  const float* drgb, const float* dpn_in, float* scratch, float* dfeat, float* dbase_params, float* dhead_params,
                                  float* dembedding, const float* directions, float* ddirections) {
    NVO_CHECK(B >= 0 && S >= 32, "field_backward: bad shape B=%lld S=%d (at least 32 samples per ray: a warp's rows span at most two rays)", (long long)B, S);
    if (B == 0) return 0;
    NVO_CHECK(feat16 && saved && wimage && rgb && h0 && selector && drgb && scratch && dfeat && dbase_params && dhead_params, "field_backward: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = B * S;
    cudaError_t e = cudaMemsetAsync(scratch, 0, 2 * sizeof(float), st);
    NVO_CHECK(e == cudaSuccess, "field_backward: memset: %s", cudaGetErrorString(e));
    k_field_bwd_absmax<<<(unsigned int)min((int64_t)nvo_sm_count() * 4, (n + 255) / 256), 256, 0, st>>>(n, drgb, rgb, ddensity, selector, h0, scratch);
    NVO_CUDA_LAUNCH_CHECK("field_backward(absmax)");
    FieldBwdP p;
    p.n = n, p.S = S, p.saved_chunks = save_pn ? FS_CHUNKS : FS_CHUNKS_HEAD, p.feat16 = (const unsigned char*)feat16, p.saved = (const unsigned char*)saved;
    p.wimg = (const unsigned char*)wimage, p.rgb = rgb, p.h0 = h0, p.sel = selector, p.cam = cam_idx, p.ddensity = ddensity, p.drgb = drgb, p.dpn_in = dpn_in;
    p.absmax = scratch, p.dfeat = dfeat, p.dbase = dbase_params, p.dhead = dhead_params, p.demb = dembedding;
    NVO_CHECK(!ddirections || directions, "field_backward: directions required for their gradient");
    p.dirs = directions, p.ddirs = ddirections;
    static const int l2_prefetch = nvo_env_int("NVO_FIELD_BWD_PREFETCH", 1);
    p.l2_prefetch = l2_prefetch;
    const int G = field_groups();
    const size_t smem = BW_GROUPS_OFF + (size_t)(G * BW_GROUP_CHUNKS + 2) * CHUNK_B + 8 * (1 + 5 * G) + 16;
    const int64_t tiles = (n + TM - 1) / TM;
    const unsigned int grid = (unsigned int)min((int64_t)nvo_sm_count(), (tiles + G - 1) / G);
    if (G == 2) {
        e = cudaFuncSetAttribute(k_field_bwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        NVO_CHECK(e == cudaSuccess, "field_backward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        k_field_bwd<2><<<grid, 2 * FT_THREADS + 64, smem, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(k_field_bwd<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        NVO_CHECK(e == cudaSuccess, "field_backward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        k_field_bwd<3><<<grid, 3 * FT_THREADS + 64, smem, st>>>(p);
    }
    NVO_CUDA_LAUNCH_CHECK("field_backward");
    return 0;
}
