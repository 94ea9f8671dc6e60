// Standalone probe (not product code): pins the tcgen05 shared-memory descriptor conventions the MLP kernels rely on.
//   T1  D[128x64] = A[128x64] * W^T           A K-major, W[out][in] K-major                    (forward layer)
//   T2  D[128x64] = G[128x64] * W             G K-major, same W smem read MN-major             (dgrad)
//   T3  D[64(pad 128)x64] = G^T * A           both operands = the natural row tiles read MN-major, K = 128 samples (wgrad)
// Each test is run for the candidate (LBO,SBO) assignments and prints the max abs error vs a CPU fp32 reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e = (x);                                                                    \
        if (e != cudaSuccess) {                                                                 \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);      \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version for sm_100
    return d;                // swizzle = 0 (none), base offset 0
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                       // D format f32
    d |= 0u << 7;                       // A f16
    d |= 0u << 10;                      // B f16
    d |= (uint32_t)a_mn_major << 15;    // 0 = K-major
    d |= (uint32_t)b_mn_major << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// natural tile layout: T[c = col/8][row][8 halfs], rows = 128 (or n_rows), i.e. chunk stride = n_rows*16 B
__device__ __forceinline__ int nat_index(int row, int col, int n_rows) { return ((col >> 3) * n_rows + row) * 8 + (col & 7); }

struct Variant {
    int test;  // 1,2,3
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    int a_mn, b_mn;
};

__global__ void __launch_bounds__(128) probe(const __half* __restrict__ A, const __half* __restrict__ G, const __half* __restrict__ W, Variant v,
                                             float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char dyn[];
    __half* sA = reinterpret_cast<__half*>(dyn);              // natural tile of A (128 rows x 64) + zero padding read by M=128 in T3
    __half* sG = sA + 128 * 64 * 2;
    __half* sW = sG + 128 * 64 * 2;                           // W[out=64][in=64], natural layout with n_rows = 64
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 64 * 2; i += 128) {
        sA[i] = __float2half(0.f);
        sG[i] = __float2half(0.f);
    }
    __syncthreads();
    for (int i = tid; i < 128 * 64; i += 128) {
        const int r = i / 64, c = i % 64;
        sA[nat_index(r, c, 128)] = A[i];
        sG[nat_index(r, c, 128)] = G[i];
    }
    for (int i = tid; i < 64 * 64; i += 128) {
        const int r = i / 64, c = i % 64;
        sW[nat_index(r, c, 64)] = W[i];
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the async proxy (UMMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const int M = 128, N = 64;
        const uint32_t idesc = make_idesc(M, N, v.a_mn, v.b_mn);
        const __half* a_src = v.test == 1 ? sA : sG;
        const __half* b_src = v.test == 3 ? sA : sW;
        const int ksteps = v.test == 3 ? 8 : 4;  // K = 128 samples or 64 features, 16 per instruction
        for (int k = 0; k < ksteps; ++k) {
            // advancing K by 16 elements: K-major -> 2 chunks (2*chunk stride); MN-major (K = rows of the natural tile) -> 16 rows * 16 B
            uint32_t a_off, b_off;
            if (v.test == 1) { a_off = k * 2 * 128 * 16; b_off = k * 2 * 64 * 16; }
            else if (v.test == 2) { a_off = k * 2 * 128 * 16; b_off = k * 16 * 16; }
            else { a_off = k * 16 * 16; b_off = k * 16 * 16; }
            const uint64_t ad = make_desc(smem_u32(a_src) + a_off, v.a_lbo, v.a_sbo);
            const uint64_t bd = make_desc(smem_u32(b_src) + b_off, v.b_lbo, v.b_sbo);
            mma_f16(tmem, ad, bd, idesc, k > 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    // wait for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(smem_u32(&mbar)), "r"(0)
                : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // each warp reads its 32 lanes, 64 columns
    uint32_t r[64];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c = 0; c < 64; c += 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[c]), "=r"(r[c + 1]), "=r"(r[c + 2]), "=r"(r[c + 3]), "=r"(r[c + 4]), "=r"(r[c + 5]), "=r"(r[c + 6]), "=r"(r[c + 7])
                     : "r"(taddr + c));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 64; ++c) out[tid * 64 + c] = __uint_as_float(r[c]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    const int M = 128, K = 64, N = 64;
    __half *hA = (__half*)malloc(M * K * 2), *hG = (__half*)malloc(M * N * 2), *hW = (__half*)malloc(N * K * 2);
    float *fA = (float*)malloc(M * K * 4), *fG = (float*)malloc(M * N * 4), *fW = (float*)malloc(N * K * 4);
    srand(1);
    for (int i = 0; i < M * K; ++i) { hA[i] = __float2half((rand() % 2001 - 1000) / 1000.f); fA[i] = __half2float(hA[i]); }
    for (int i = 0; i < M * N; ++i) { hG[i] = __float2half((rand() % 2001 - 1000) / 1000.f); fG[i] = __half2float(hG[i]); }
    for (int i = 0; i < N * K; ++i) { hW[i] = __float2half((rand() % 2001 - 1000) / 1000.f); fW[i] = __half2float(hW[i]); }
    // references
    float *r1 = (float*)calloc(M * N, 4), *r2 = (float*)calloc(M * K, 4), *r3 = (float*)calloc(N * K, 4);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0;
            for (int k = 0; k < K; ++k) s += fA[m * K + k] * fW[n * K + k];
            r1[m * N + n] = s;  // A W^T
        }
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k) {
            float s = 0;
            for (int n = 0; n < N; ++n) s += fG[m * N + n] * fW[n * K + k];
            r2[m * K + k] = s;  // G W
        }
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            float s = 0;
            for (int m = 0; m < M; ++m) s += fG[m * N + n] * fA[m * K + k];
            r3[n * K + k] = s;  // G^T A
        }
    __half *dA, *dG, *dW;
    float* dOut;
    CK(cudaMalloc(&dA, M * K * 2));
    CK(cudaMalloc(&dG, M * N * 2));
    CK(cudaMalloc(&dW, N * K * 2));
    CK(cudaMalloc(&dOut, 128 * 64 * 4));
    CK(cudaMemcpy(dA, hA, M * K * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dG, hG, M * N * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, hW, N * K * 2, cudaMemcpyHostToDevice));
    float* hOut = (float*)malloc(128 * 64 * 4);
    const int SMEM = (128 * 64 * 2 * 2 + 64 * 64) * 2;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    // chunk strides of the natural layouts: 128-row tiles 2048 B, 64-row W 1024 B; 8-row group 128 B
    Variant vs[] = {
        {1, 2048, 128, 1024, 128, 0, 0},  // K-major: LBO = K-chunk stride, SBO = 8-row group stride
        {1, 128, 2048, 128, 1024, 0, 0},  // swapped
        {2, 2048, 128, 128, 1024, 0, 1},  // B MN-major: LBO = 8-k group stride (128 B), SBO = mn-chunk stride
        {2, 2048, 128, 1024, 128, 0, 1},  // swapped
        {3, 128, 2048, 128, 2048, 1, 1},  // both MN-major
        {3, 2048, 128, 2048, 128, 1, 1},  // swapped
    };
    int vi = -1;
    for (auto& v : vs) {
        ++vi;
        if (only >= 0 && vi != only) continue;
        CK(cudaMemset(dOut, 0, 128 * 64 * 4));
        probe<<<1, 128, SMEM>>>(dA, dG, dW, v, dOut);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("test %d variant (a %u/%u b %u/%u): CUDA error %s\n", v.test, v.a_lbo, v.a_sbo, v.b_lbo, v.b_sbo, cudaGetErrorString(e));
            return 1;
        }
        CK(cudaMemcpy(hOut, dOut, 128 * 64 * 4, cudaMemcpyDeviceToHost));
        const float* ref = v.test == 1 ? r1 : (v.test == 2 ? r2 : r3);
        const int rows = v.test == 3 ? 64 : 128;
        float err = 0, mx = 0;
        for (int i = 0; i < rows * 64; ++i) {
            err = fmaxf(err, fabsf(hOut[i] - ref[i]));
            mx = fmaxf(mx, fabsf(ref[i]));
        }
        printf("test %d  A(lbo=%u,sbo=%u,mn=%d) B(lbo=%u,sbo=%u,mn=%d): max abs err %.5f (ref max %.3f) %s\n", v.test, v.a_lbo, v.a_sbo, v.a_mn, v.b_lbo,
               v.b_sbo, v.b_mn, err, mx, err < 1e-2 ? "MATCH" : "mismatch");
    }
    return 0;
}
