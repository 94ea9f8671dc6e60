// Fused dense Adam over one flat fp32 parameter buffer: torch.optim.Adam semantics
// (NS/engine/optimizers.py:138-150 with lr 1e-2, eps 1e-15, no weight decay; nerf_vo/mapping/nerfstudio.py:84-100),
// gradient pre-scale (1/world_size for the allreduce mean) folded in.  The step counter lives on the device so the
// launch can be replayed from a CUDA graph.
#include "nvo_common.cuh"

__global__ void __launch_bounds__(256) k_adam(int64_t n4, int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, const int* __restrict__ step_ptr, float lr, float b1, float b2, float eps,
                                              float grad_scale) {
    __shared__ float s_c[2];
    if (threadIdx.x == 0) {
        const float t = (float)(*step_ptr + 1);
        s_c[0] = lr / (1.f - powf(b1, t));         // step_size
        s_c[1] = 1.f / sqrtf(1.f - powf(b2, t));   // 1/sqrt(bias_correction2)
    }
    __syncthreads();
    const float step_size = s_c[0], inv_bc2 = s_c[1];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
#define UPD(c)                                                        \
    {                                                                 \
        const float gr = gg.c * grad_scale;                           \
        mm.c = b1 * mm.c + (1.f - b1) * gr;                           \
        vv.c = b2 * vv.c + (1.f - b2) * gr * gr;                      \
        pp.c -= step_size * mm.c / (sqrtf(vv.c) * inv_bc2 + eps);     \
    }
        UPD(x) UPD(y) UPD(z) UPD(w)
#undef UPD
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    // tail (n not a multiple of 4)
    if (blockIdx.x == 0) {
        const int64_t j = n4 * 4 + threadIdx.x;
        if (j < n) {
            const float gr = g[j] * grad_scale;
            const float mj = b1 * m[j] + (1.f - b1) * gr;
            const float vj = b2 * v[j] + (1.f - b2) * gr * gr;
            m[j] = mj;
            v[j] = vj;
            p[j] -= step_size * mj / (sqrtf(vj) * inv_bc2 + eps);
        }
    }
}

__global__ void k_tick(int* step) { *step += 1; }

extern "C" int nvo_adam_step(void* stream, int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int32_t* step, float lr,
                             float beta1, float beta2, float eps, float grad_scale) {
    NVO_CHECK(n >= 0, "adam_step: negative size");
    if (n == 0) return 0;
    NVO_CHECK(params && grads && exp_avg && exp_avg_sq && step, "adam_step: null pointer");
    NVO_CHECK(((uintptr_t)params % 16 == 0) && ((uintptr_t)grads % 16 == 0) && ((uintptr_t)exp_avg % 16 == 0) && ((uintptr_t)exp_avg_sq % 16 == 0),
              "adam_step: buffers must be 16-byte aligned");
    const int64_t n4 = n / 4;
    cudaStream_t st = (cudaStream_t)stream;
    // largest shared-memory carve-out: kernels with dynamic shared memory (proposal backward) can then share an SM with this one
    // (the L1 / shared split of an SM cannot change while CTAs are resident); Adam streams and has no use for L1
    cudaFuncSetAttribute(k_adam, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    k_adam<<<nvo_blocks(n4 > 0 ? n4 : 1, 256), 256, 0, st>>>(n4, n, params, grads, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, grad_scale);
    NVO_CUDA_LAUNCH_CHECK("adam_step");
    k_tick<<<1, 1, 0, st>>>(step);
    NVO_CUDA_LAUNCH_CHECK("adam_step(tick)");
    return 0;
}
