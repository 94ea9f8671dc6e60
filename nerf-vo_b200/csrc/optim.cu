// Fused dense Adam over one flat fp32 parameter buffer: torch.optim.Adam semantics
// (NS/engine/optimizers.py:138-150 with lr 1e-2, eps 1e-15, no weight decay; nerf_vo/mapping/nerfstudio.py:84-100),
// gradient pre-scale (1/world_size for the allreduce mean) folded in.  The step counter lives on the device so the
// launch can be replayed from a CUDA graph.
#include "nvo_common.cuh"
#include "adam.cuh"

__global__ void __launch_bounds__(256) k_adam(int64_t n4, int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, const int* __restrict__ step_ptr, double lr, double b1, double b2, double eps,
                                              float grad_scale) {
    __shared__ AdamC s_c;
    if (threadIdx.x == 0) s_c = adam_constants(*step_ptr + 1, lr, b1, b2, eps, grad_scale);
    __syncthreads();
    const AdamC c = s_c;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        adam_update(pp.x, mm.x, vv.x, gg.x, c);
        adam_update(pp.y, mm.y, vv.y, gg.y, c);
        adam_update(pp.z, mm.z, vv.z, gg.z, c);
        adam_update(pp.w, mm.w, vv.w, gg.w, c);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    // tail (n not a multiple of 4)
    if (blockIdx.x == 0) {
        const int64_t j = n4 * 4 + threadIdx.x;
        if (j < n) {
            float pj = p[j], mj = m[j], vj = v[j];
            adam_update(pj, mj, vj, g[j], c);
            p[j] = pj, m[j] = mj, v[j] = vj;
        }
    }
}

__global__ void k_tick(int* step) { *step += 1; }

extern "C" int nvo_adam_step(void* stream, int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int32_t* step, double lr,
                             double beta1, double beta2, double eps, float grad_scale) {
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
