// Standalone probe (not product code): issue rate of tcgen05.mma kind::f16, cta_group::1, M=128 with the no-swizzle canonical
// shared-memory layout the MLP kernels use, for the three operand arrangements (forward K-major/K-major, dgrad K-major/MN-major,
// wgrad MN-major/MN-major) and several N.  One CTA per SM, R back-to-back MMAs into one accumulator, clock64 around issue..commit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate umma_rate.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t swz) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)swz << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(acc)
                 : "memory");
}

// mode 0: fwd (A K-major lbo 2048/sbo 128, B K-major lbo N*16/sbo 128); 1: dgrad (B MN-major lbo 128/sbo N*16); 2: wgrad (both MN-major lbo 128/sbo 2048)
// mode 3: fwd with 128B-swizzle descriptors (K-major, sbo 1024) — data is garbage, only the rate matters
// mode 4 / 5: as mode 0 (no swizzle) / mode 3 (128B swizzle) but every MMA reads DIFFERENT operand tiles (four A tiles and four B tiles in
// rotation, as the K steps of a layer do): the same-descriptor loops above may be served from an operand cache
__global__ void __launch_bounds__(128) rate(int mode, int N, int R, int two_acc, long long* out) {
    extern __shared__ __align__(1024) unsigned char dyn[];
    __shared__ __align__(8) uint64_t mbar, mbar2;
    __shared__ volatile int stop;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(dyn)[i] = 0x3C003C00u;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar2)) : "memory");
        stop = 0;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t a0 = smem_u32(dyn), b0 = smem_u32(dyn + 32768);
        uint32_t idesc;
        uint64_t ad, bd;
        if (mode == 0) { idesc = make_idesc(128, N, 0, 0); ad = make_desc(a0, 2048, 128, 0); bd = make_desc(b0, N * 16, 128, 0); }
        else if (mode == 1) { idesc = make_idesc(128, N, 0, 1); ad = make_desc(a0, 2048, 128, 0); bd = make_desc(b0, 128, N * 16, 0); }
        else if (mode == 2 || mode == 6 || mode == 7) { idesc = make_idesc(128, N, 1, 1); ad = make_desc(a0, 128, 2048, 0); bd = make_desc(b0, 128, 2048, 0); }
        else if (mode == 3 || mode == 5) { idesc = make_idesc(128, N, 0, 0); ad = make_desc(a0, 16, 1024, 2); bd = make_desc(b0, 16, 1024, 2); }
        else { idesc = make_idesc(128, N, 0, 0); ad = make_desc(a0, 2048, 128, 0); bd = make_desc(b0, N * 16, 128, 0); }
        t0 = clock64();
        if (mode == 6 || mode == 7) {
            // wgrad as the field backward issues it: 8 K-steps of 16 samples, A and B slabs 256 B apart, four different tile pairs in rotation
            for (int r = 0; r < R; r += 8) {
                const uint64_t o = (uint64_t)(((r >> 3) & 3) * 2048) >> 4;
#pragma unroll
                for (int k = 0; k < 8; ++k) mma_f16(tmem, ad + o + (uint64_t)k * 16, bd + o + (uint64_t)k * 16, idesc, (r | k) > 0);
            }
        } else if (mode == 8) {
            // the forward's pattern: a burst of `two_acc ? 5 : 12` MMAs, commit, wait — latency of a small batch, not throughput
            const int burst = two_acc ? 5 : 12;
            uint32_t ph = 0;
            for (int r = 0; r < R; r += burst) {
                for (int k = 0; k < burst; ++k) mma_f16(tmem, ad + (uint64_t)(k & 3) * 256, bd + (uint64_t)(k & 3) * (N * 2), idesc, k > 0);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar2)) : "memory");
                uint32_t dn = 0;
                while (!dn)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(dn) : "r"(smem_u32(&mbar2)), "r"(ph) : "memory");
                ph ^= 1;
            }
        } else if (mode >= 4) {
            // operand tiles 4 KB (A: 128 rows x 16 columns) / N * 32 B (B) apart: descriptor start address field += bytes >> 4
            const uint64_t a_step = 4096 >> 4, b_step = (uint64_t)(N * 32) >> 4;
            for (int r = 0; r < R; r += 4) {
                mma_f16(tmem, ad, bd, idesc, r > 1);
                mma_f16(tmem, ad + a_step, bd + b_step, idesc, 1);
                mma_f16(tmem, ad + 2 * a_step, bd + 2 * b_step, idesc, 1);
                mma_f16(tmem, ad + 3 * a_step, bd + 3 * b_step, idesc, 1);
            }
        } else
        for (int r = 0; r < R; ++r) mma_f16(tmem + (two_acc ? (r & 1) * 256 : 0), ad, bd, idesc, r > 1);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    if (mode == 7 && warp >= 1) {
        uint4* z = reinterpret_cast<uint4*>(dyn + 49152) + (tid - 32);  // a region no MMA reads
        uint4 v = make_uint4(tid, 0, 0, 0);
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                z[(i & 3) * 256] = v;
                v.x += z[((i + 1) & 3) * 256].y;
            }
        }
        if (v.x == 0x12345678u) out[0] = 0;
    }
    uint32_t done = 0;
    if (!(mode == 7 && warp >= 1))
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
    if (tid == 0) {
        t1 = clock64();
        out[blockIdx.x] = t1 - t0;
        stop = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * 8);
    long long h[148];
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int R = 4080;  // a multiple of 8, 12 and 5
    const char* names[] = {"fwd  K/K  no-swizzle", "dgrad K/MN no-swizzle", "wgrad MN/MN no-swizzle", "fwd  K/K  128B-swizzle",
                           "fwd no-swizzle, rotating tiles", "fwd 128B-swizzle, rotating tiles", "wgrad MN/MN rotating slabs", "wgrad rotating + smem noise",
                           "fwd bursts (acc=1: 12, acc=2: 5 MMAs) + commit + wait"};
    const int first_mode = getenv("UMMA_RATE_FIRST") ? atoi(getenv("UMMA_RATE_FIRST")) : 0;
    for (int mode = first_mode; mode < 9; ++mode)
        for (int N : {16, 32, 64, 128, 256})
            for (int two = 0; two < 2; ++two) {
                if ((mode == 1 || mode == 2) && N > 128) continue;
                if (mode >= 4 && mode != 8 && (two || N > 128)) continue;
                if (mode >= 6 && N > 64) continue;
                rate<<<148, 128, 65536>>>(mode, N, R, two, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s N=%d: %s\n", names[mode], N, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                double s = 0;
                for (int i = 0; i < 148; ++i) s += h[i];
                printf("%-24s N=%3d acc=%d : %.1f cycles/MMA (M=128,K=16)  floor=%.0f\n", names[mode], N, two + 1, s / 148 / R, 128.0 * N / 256);
            }
    return 0;
}
