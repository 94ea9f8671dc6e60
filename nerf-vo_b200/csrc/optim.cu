// Fused dense Adam over one flat fp32 parameter buffer: torch.optim.Adam semantics
// (NS/engine/optimizers.py:138-150 with lr 1e-2, eps 1e-15, no weight decay; nerf_vo/mapping/nerfstudio.py:84-100),
// gradient pre-scale (1/world_size for the allreduce mean) folded in.  The step counter lives on the device so the
// launch can be replayed from a CUDA graph.
#include "nvo_common.cuh"
#include "adam.cuh"

// Persistent grid (a few CTAs per SM, grid-stride over float4 items): the Adam constants cost a double-precision pow() per CTA, which a
// one-item-per-thread launch pays once per WAVE (16 k CTAs = 14 waves of ~3 us of exposed latency each: measured 124 us = 58 % of the HBM
// peak for the 16.8 M-parameter group); here it is paid once, behind the first batch of loads, and every thread keeps 4 x 4 float4 loads in flight.
#define ADAM_UNROLL 4
__global__ void __launch_bounds__(256) k_adam(int64_t n4, int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, const int* __restrict__ step_ptr, double lr, double b1, double b2, double eps,
                                              float grad_scale, double lr_final, int max_steps) {
    __shared__ AdamC s_c;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // first batch of loads in flight before the constants are needed
    float4 pp[ADAM_UNROLL], gg[ADAM_UNROLL], mm[ADAM_UNROLL], vv[ADAM_UNROLL];
#pragma unroll
    for (int u = 0; u < ADAM_UNROLL; ++u) {
        const int64_t j = i + u * stride;
        if (j < n4) {
            pp[u] = reinterpret_cast<float4*>(p)[j];
            gg[u] = __ldg(reinterpret_cast<const float4*>(g) + j);
            mm[u] = reinterpret_cast<float4*>(m)[j];
            vv[u] = reinterpret_cast<float4*>(v)[j];
        }
    }
    if (threadIdx.x == 0) {
        const int done = *step_ptr;  // optimizer steps taken so far == the epoch the LambdaLR scheduler is in
        if (lr_final > 0.0) {
            // ExponentialDecayScheduler without warm-up (NS/engine/schedulers.py:122-138): exp(log(lr_init) (1 - t) + log(lr_final) t)
            const double t = fmin(fmax((double)done / (double)max_steps, 0.0), 1.0);
            lr = exp(log(lr) * (1.0 - t) + log(lr_final) * t);
        }
        s_c = adam_constants(done + 1, lr, b1, b2, eps, grad_scale);
    }
    __syncthreads();
    const AdamC c = s_c;
    while (i < n4) {
#pragma unroll
        for (int u = 0; u < ADAM_UNROLL; ++u) {
            const int64_t j = i + u * stride;
            if (j < n4) {
                adam_update(pp[u].x, mm[u].x, vv[u].x, gg[u].x, c);
                adam_update(pp[u].y, mm[u].y, vv[u].y, gg[u].y, c);
                adam_update(pp[u].z, mm[u].z, vv[u].z, gg[u].z, c);
                adam_update(pp[u].w, mm[u].w, vv[u].w, gg[u].w, c);
                reinterpret_cast<float4*>(p)[j] = pp[u];
                reinterpret_cast<float4*>(m)[j] = mm[u];
                reinterpret_cast<float4*>(v)[j] = vv[u];
            }
        }
        i += ADAM_UNROLL * stride;
#pragma unroll
        for (int u = 0; u < ADAM_UNROLL; ++u) {
            const int64_t j = i + u * stride;
            if (j < n4) {
                pp[u] = reinterpret_cast<float4*>(p)[j];
                gg[u] = __ldg(reinterpret_cast<const float4*>(g) + j);
                mm[u] = reinterpret_cast<float4*>(m)[j];
                vv[u] = reinterpret_cast<float4*>(v)[j];
            }
        }
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

// one float4 per thread, constants evaluated by thread 0 of every CTA (the round-1 kernel; NVO_ADAM_MODE=0) or read from `consts` when
// a one-thread pre-kernel evaluated them (NVO_ADAM_MODE=2).  Six CTAs per SM (40 registers instead of 43: 48 instead of 40 resident warps,
// 109 -> 104 us L2-flushed); eight (32 registers) spills and is slower.
__global__ void __launch_bounds__(256, 6) k_adam_flat(int64_t n4, int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, const int* __restrict__ step_ptr, double lr, double b1, double b2, double eps,
                                                   float grad_scale, double lr_final, int max_steps, const AdamC* __restrict__ consts) {
    __shared__ AdamC s_c;
    if (threadIdx.x == 0) {
        if (consts) {
            s_c = *consts;
        } else {
            const int done = *step_ptr;
            if (lr_final > 0.0) {
                const double t = fmin(fmax((double)done / (double)max_steps, 0.0), 1.0);
                lr = exp(log(lr) * (1.0 - t) + log(lr_final) * t);
            }
            s_c = adam_constants(done + 1, lr, b1, b2, eps, grad_scale);
        }
    }
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
    if (blockIdx.x == 0) {
        const int64_t j = n4 * 4 + threadIdx.x;
        if (j < n) {
            float pj = p[j], mj = m[j], vj = v[j];
            adam_update(pj, mj, vj, g[j], c);
            p[j] = pj, m[j] = mj, v[j] = vj;
        }
    }
}
__device__ AdamC g_adam_consts[64];
__global__ void k_adam_consts(const int* __restrict__ step_ptr, double lr, double b1, double b2, double eps, float grad_scale, double lr_final, int max_steps,
                              AdamC* __restrict__ out) {
    const int done = *step_ptr;
    if (lr_final > 0.0) {
        const double t = fmin(fmax((double)done / (double)max_steps, 0.0), 1.0);
        lr = exp(log(lr) * (1.0 - t) + log(lr_final) * t);
    }
    *out = adam_constants(done + 1, lr, b1, b2, eps, grad_scale);
}

__global__ void k_tick(int* step) { *step += 1; }

static int adam_step_impl(void* stream, int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int32_t* step, double lr,
                          double beta1, double beta2, double eps, float grad_scale, double lr_final, int max_steps) {
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
    static const int mode = nvo_env_int("NVO_ADAM_MODE", 2);  // 0: constants per CTA, 1: persistent grid, 2: pre-kernel constants (default)
    if (mode == 1) {
        const int64_t want = (n4 + 256 * ADAM_UNROLL - 1) / (256 * ADAM_UNROLL);
        const unsigned int grid = (unsigned int)max((int64_t)1, min((int64_t)nvo_sm_count() * 2, want));
        k_adam<<<grid, 256, 0, st>>>(n4, n, params, grads, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, grad_scale, lr_final, max_steps);
    } else {
        // the constants (a double-precision pow per launch) come from a one-thread pre-kernel: evaluated by thread 0 of every CTA they cost
        // every WAVE of the 16 k-CTA launch a few microseconds of exposed latency (measured: step 0.812 -> 0.800 ms).  64 device slots used
        // round-robin, so launches in flight on different streams (the two parameter groups) never share one.
        AdamC* consts = nullptr;
        if (mode == 2) {
            static int next = 0;
            AdamC* slots = nullptr;
            cudaGetSymbolAddress((void**)&slots, g_adam_consts);
            consts = slots + (next++ & 63);
            k_adam_consts<<<1, 1, 0, st>>>(step, lr, beta1, beta2, eps, grad_scale, lr_final, max_steps, consts);
            NVO_CUDA_LAUNCH_CHECK("adam_step(consts)");
        }
        cudaFuncSetAttribute(k_adam_flat, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        // NVO_ADAM_SMEM bytes of (unused) dynamic shared memory per CTA cap the CTAs per SM, leaving thread slots to kernels running next to it
        static const int pad_smem = nvo_env_int("NVO_ADAM_SMEM", 0);
        if (pad_smem > 0) cudaFuncSetAttribute(k_adam_flat, cudaFuncAttributeMaxDynamicSharedMemorySize, pad_smem);
        k_adam_flat<<<nvo_blocks(n4 > 0 ? n4 : 1, 256), 256, pad_smem > 0 ? pad_smem : 0, st>>>(n4, n, params, grads, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, grad_scale,
                                                                      lr_final, max_steps, consts);
    }
    NVO_CUDA_LAUNCH_CHECK("adam_step");
    k_tick<<<1, 1, 0, st>>>(step);
    NVO_CUDA_LAUNCH_CHECK("adam_step(tick)");
    return 0;
}

extern "C" int nvo_adam_step(void* stream, int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int32_t* step, double lr,
                             double beta1, double beta2, double eps, float grad_scale) {
    return adam_step_impl(stream, n, params, grads, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, grad_scale, 0.0, 1);
}

// the same with the learning rate of ExponentialDecayScheduler (no warm-up) evaluated on the device from the step counter, so a captured
// CUDA graph follows the schedule (the "camera_opt" group: lr 1e-4 -> 1e-5 over the mapping iterations, nerf_vo/mapping/nerfstudio.py:93-100)
extern "C" int nvo_adam_step_decay(void* stream, int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int32_t* step,
                                   double lr_init, double lr_final, int32_t max_steps, double beta1, double beta2, double eps, float grad_scale) {
    NVO_CHECK(lr_init > 0.0 && lr_final > 0.0 && max_steps >= 1, "adam_step_decay: bad schedule");
    return adam_step_impl(stream, n, params, grads, exp_avg, exp_avg_sq, step, lr_init, beta1, beta2, eps, grad_scale, lr_final, max_steps);
}

// ---- CameraOptimizer.get_loss_dict (NS/cameras/camera_optimizers.py:149-155): mean ||t_k|| trans_l2_penalty + mean ||r_k|| rot_l2_penalty and
// its gradient (torch's norm backward: x / ||x||, zero at x = 0) accumulated into d_pose[K,6] scaled by `scale` ---------------------------------
__global__ void __launch_bounds__(256) k_pose_regularizer(int K, const float* __restrict__ pose, float trans_pen, float rot_pen, float scale,
                                                          float* __restrict__ loss, float* __restrict__ d_pose) {
    __shared__ float red[8];
    float v = 0.f;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
        const float* q = pose + 6 * k;
        const float nt = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]), nr = sqrtf(q[3] * q[3] + q[4] * q[4] + q[5] * q[5]);
        v += (nt * trans_pen + nr * rot_pen) / (float)K;
        if (d_pose) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (nt > 0.f) d_pose[6 * k + a] += scale * trans_pen / (float)K * q[a] / nt;
                if (nr > 0.f) d_pose[6 * k + 3 + a] += scale * rot_pen / (float)K * q[3 + a] / nr;
            }
        }
    }
    v = nvo_warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0 && loss) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        atomicAdd(loss, s);
    }
}

extern "C" int nvo_pose_regularizer(void* stream, int32_t K, const float* pose_adjustment, float trans_l2_penalty, float rot_l2_penalty, float scale,
                                    float* loss, float* d_pose) {
    NVO_CHECK(K >= 1 && pose_adjustment, "pose_regularizer: bad arguments");
    k_pose_regularizer<<<1, 256, 0, (cudaStream_t)stream>>>(K, pose_adjustment, trans_l2_penalty, rot_l2_penalty, scale, loss, d_pose);
    NVO_CUDA_LAUNCH_CHECK("pose_regularizer");
    return 0;
}
