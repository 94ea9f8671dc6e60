// Standalone probe (not product code): L2 reduction throughput on B200 for the scatter patterns of the hash-grid backward.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/atomics_probe tools/atomics_probe.cu
// Each thread issues R reductions to pseudo-random rows of a table of T rows (row = 8 bytes: two floats).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// mode 0: red.f32 x2 (two scalar ops per row)   1: red.v2.f32   2: red.v4.f32 on aligned row pairs (one op per 2 rows)
// mode 3: red.f16x2 (one 4-byte op per row)     4: red.v2.f32 with lane pairs hitting adjacent rows (same 16 B)
// mode 5: red.v2.f32 where all 32 lanes of a warp hit the same 256-byte block (perfect sector locality)
template <int MODE>
__global__ void k(float* table, uint32_t mask, int R, uint32_t seed) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < R; ++r) {
        uint32_t row = hash32(t * 977u + r * 131071u + seed) & mask;
        float a = 1.0f, b = 2.0f;
        if (MODE == 0) {
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(table + 2 * (size_t)row), "f"(a) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(table + 2 * (size_t)row + 1), "f"(b) : "memory");
        } else if (MODE == 1) {
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(table + 2 * (size_t)row), "f"(a), "f"(b) : "memory");
        } else if (MODE == 2) {
            row &= ~1u;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(table + 2 * (size_t)row), "f"(a), "f"(b), "f"(a), "f"(b) : "memory");
        } else if (MODE == 3) {
            __half2 h = __floats2half2_rn(a, b);
            asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(reinterpret_cast<__half2*>(table) + row), "r"(*reinterpret_cast<uint32_t*>(&h)) : "memory");
        } else if (MODE == 4) {
            row = (hash32((t >> 1) * 977u + r * 131071u + seed) & mask & ~1u) | (t & 1u);
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(table + 2 * (size_t)row), "f"(a), "f"(b) : "memory");
        } else if (MODE == 5) {
            row = (hash32((t >> 5) * 977u + r * 131071u + seed) & mask & ~31u) | (t & 31u);
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(table + 2 * (size_t)row), "f"(a), "f"(b) : "memory");
        } else if (MODE == 6) {  // all 32 lanes of a warp on the SAME row
            row = hash32((t >> 5) * 977u + r * 131071u + seed) & mask;
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(table + 2 * (size_t)row), "f"(a), "f"(b) : "memory");
        } else if (MODE == 7) {  // groups of 8 lanes on the same (random) row: 4 random rows per warp
            row = hash32((t >> 3) * 977u + r * 131071u + seed) & mask;
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(table + 2 * (size_t)row), "f"(a), "f"(b) : "memory");
        } else {  // groups of 2 lanes on the same (random) row
            row = hash32((t >> 1) * 977u + r * 131071u + seed) & mask;
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(table + 2 * (size_t)row), "f"(a), "f"(b) : "memory");
        }
    }
}

template <int MODE>
void run(const char* name, float* table, int log2T, int threads_total, int R) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const uint32_t mask = (1u << log2T) - 1;
    k<MODE><<<threads_total / 256, 256>>>(table, mask, R, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k<MODE><<<threads_total / 256, 256>>>(table, mask, R, 7 + i);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double rows = 5.0 * threads_total * R * (MODE == 2 ? 2 : 1);
    printf("%-44s T=2^%d  %.1f us/launch  %.1f G rows/s  (%.1f G instr-lanes/s)\n", name, log2T, ms * 200, rows / (ms * 1e-3) / 1e9,
           5.0 * threads_total * R * (MODE == 0 ? 2 : 1) / (ms * 1e-3) / 1e9);
}

int main() {
    float* table;
    cudaMalloc(&table, (size_t)8 << 24);  // 2^24 rows x 8 B = 128 MiB
    cudaMemset(table, 0, (size_t)8 << 24);
    const int threads = 1 << 20, R = 16;
    for (int log2T : {17, 19}) {
        run<0>("0: 2 x red.f32 per row", table, log2T, threads, R);
        run<1>("1: red.v2.f32 per row", table, log2T, threads, R);
        run<2>("2: red.v4.f32 per aligned row pair", table, log2T, threads, R);
        run<3>("3: red.f16x2 per row", table, log2T, threads, R);
        run<4>("4: red.v2.f32, lane pairs on adjacent rows", table, log2T, threads, R);
        run<5>("5: red.v2.f32, warp on one 256 B block", table, log2T, threads, R);
        run<6>("6: red.v2.f32, 32 lanes same row", table, log2T, threads, R);
        run<7>("7: red.v2.f32, 8 lanes per row", table, log2T, threads, R);
        run<8>("8: red.v2.f32, 2 lanes per row", table, log2T, threads, R);
    }
    return 0;
}
