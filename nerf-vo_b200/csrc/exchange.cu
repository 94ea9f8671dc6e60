// Data-parallel gradient exchange fused with the optimizer, over NVLink / NVSwitch peer memory (SURVEY §8e + §8f row 1).
//
// The reference would wrap the model in DistributedDataParallel (NS/pipelines/base_pipeline.py:281-283: bucketed NCCL
// all-reduce of every gradient, then a dense torch Adam on every rank).  Here the exchange and the optimizer are ONE kernel:
//
//   rank r owns the slice [r*chunk, (r+1)*chunk) of the flat parameter buffer (ZeRO-1 style: its Adam moments exist only on r)
//     1. barrier "gradients ready": release-store of the step epoch into every peer's flag pad, acquire-poll of the own pad;
//     2. reduce-scatter by direct peer LOADS: g = sum over ranks of grad_k[i] for i in the own slice (16-byte loads through NVLink);
//     3. Adam on the slice (moments local, 1/W of the optimizer traffic of a replicated Adam);
//     4. all-gather by direct peer STORES: the new parameter values are written into every rank's replica;
//     5. barrier "replicas written": system fence, the last CTA signals every peer and waits for theirs — after the kernel a rank
//        may zero its gradient buffer and read its parameters.
//
// Per step and GPU, (W-1)/W of the flat buffer crosses NVLink once in each direction (the minimum for an all-reduce), no
// staging copies, no separate optimizer pass; the whole mapping step stays one CUDA graph (no host-side collective call).
//
// Buffers that peers touch (parameters, gradients, flag pads) are plain cudaMalloc allocations exported with CUDA IPC
// (nvo_peer_alloc / nvo_peer_open); PyTorch wraps them as tensors.  Spin waits are bounded (~4 s): on timeout the kernel
// raises the error word instead of hanging the GPU.
#include "nvo_common.cuh"
#include "adam.cuh"
#include <stdlib.h>

#define NVO_MAX_PEERS 16
#define FLAG_READY 0                 // flags[FLAG_READY + k]: rank k's gradients of epoch e are complete
#define FLAG_DONE NVO_MAX_PEERS      // flags[FLAG_DONE + k]:  rank k has finished reading / writing peers for epoch e
#define FLAG_COUNT (2 * NVO_MAX_PEERS)   // flags[FLAG_COUNT]: local CTA completion counter; [FLAG_COUNT+1]: error word
// one block of flags per exchange PHASE (parameter group): the groups' exchanges of one step run as separate launches that may be in
// flight at the same time (the fields group is exchanged while the proposal networks' backward still runs)
#define FLAG_PHASE_STRIDE 40
#define NVO_MAX_PHASES 3
#define FLAG_WORDS 128

struct PeerSet {
    float* params[NVO_MAX_PEERS];
    const float* grads[NVO_MAX_PEERS];
    int* flags[NVO_MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
    float4 v;  // system-scope relaxed load: never served from a stale L1 line of a previous step
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer_f4(float* p, const float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// waits until flags[base + k] >= epoch for every k != rank (thread k polls flag k); false on timeout
__device__ __forceinline__ bool wait_all(const int* flags, int base, int world, int rank, int epoch) {
    bool ok = true;
    const int k = threadIdx.x;
    if (k < world && k != rank) {
        const long long t0 = clock64();
        while (ld_acquire_sys(flags + base + k) < epoch) {
            if (clock64() - t0 > 8000000000LL) {  // ~4 s at 2 GHz
                ok = false;
                break;
            }
            __nanosleep(64);
        }
    }
    return __syncthreads_and(ok);
}

// U float4 items per thread and trip: U*W independent 16-byte peer loads in flight per thread (NVLink round trips are ~2 us; the
// kernel is latency-bound unless every SM keeps tens of KB outstanding)
struct SliceArgs {
    int64_t lo4, hi4;   // the rank's slice of one parameter group, in float4 units of the flat buffers
    float* m;           // Adam moments of the slice (index 0 = element lo4)
    float* v;
    const int* step;    // the group's step counter
};

// reduce-scatter (peer loads) -> Adam -> all-gather (peer stores) on one slice
template <int W, int U>
__device__ __forceinline__ void slice_pass(const PeerSet& ps, int rank, const SliceArgs& a, const AdamC& c) {
    if (a.hi4 <= a.lo4) return;
    float* __restrict__ m = a.m;
    float* __restrict__ v = a.v;
    const int64_t lo4 = a.lo4, hi4 = a.hi4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi4; i0 += stride * U) {
        float4 g[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < hi4) {
#pragma unroll
                for (int k = 0; k < W; ++k) g[u][k] = ld_peer_f4(ps.grads[k] + 4 * i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i >= hi4) break;
            float4 gs = g[u][0];
#pragma unroll
            for (int k = 1; k < W; ++k) gs.x += g[u][k].x, gs.y += g[u][k].y, gs.z += g[u][k].z, gs.w += g[u][k].w;
            const int64_t j = i - lo4;
            float4 pp = reinterpret_cast<const float4*>(ps.params[rank])[i];
            float4 mm = reinterpret_cast<float4*>(m)[j];
            float4 vv = reinterpret_cast<float4*>(v)[j];
            adam_update(pp.x, mm.x, vv.x, gs.x, c);
            adam_update(pp.y, mm.y, vv.y, gs.y, c);
            adam_update(pp.z, mm.z, vv.z, gs.z, c);
            adam_update(pp.w, mm.w, vv.w, gs.w, c);
            reinterpret_cast<float4*>(m)[j] = mm;
            reinterpret_cast<float4*>(v)[j] = vv;
#pragma unroll
            for (int k = 0; k < W; ++k) st_peer_f4(ps.params[k] + 4 * i, pp);
        }
    }
}

// One launch = barrier, the rank's slice of group A (and, when present, of group B: two parameter groups stepped together share the two
// barriers), barrier.  The barrier epoch is group A's step counter.
template <int W, int U>
__global__ void __launch_bounds__(256) k_exchange_adam(const __grid_constant__ PeerSet ps, int rank, const __grid_constant__ SliceArgs sa,
                                                       const __grid_constant__ SliceArgs sb, double lr, double b1, double b2, double eps,
                                                       float grad_scale, int flag_base, double lr_final, int max_steps) {
    int* my_flags = ps.flags[rank] + flag_base;
    const int epoch = *sa.step + 1;
    // ---- 1. every rank's backward has landed -----------------------------------------------------------------------
    if (blockIdx.x == 0 && threadIdx.x < W && threadIdx.x != rank) {
        __threadfence_system();
        st_release_sys(ps.flags[threadIdx.x] + flag_base + FLAG_READY + rank, epoch);
    }
    if (!wait_all(my_flags, FLAG_READY, W, rank, epoch)) {
        if (threadIdx.x == 0) atomicExch(my_flags + FLAG_COUNT + 1, 1);
    }
    // ---- 2-4. reduce-scatter (peer loads) -> Adam -> all-gather (peer stores) on the own slices -------------------------
    // Adam constants per group (each has its own step counter), evaluated once per CTA in double precision
    __shared__ AdamC s_adam[2];
    if (threadIdx.x == 0) {
        double lr_a = lr;
        if (lr_final > 0.0) {  // group A under ExponentialDecayScheduler (the "camera_opt" group), see k_adam
            const double t = fmin(fmax((double)*sa.step / (double)max_steps, 0.0), 1.0);
            lr_a = exp(log(lr) * (1.0 - t) + log(lr_final) * t);
        }
        s_adam[0] = adam_constants(*sa.step + 1, lr_a, b1, b2, eps, grad_scale);
        s_adam[1] = adam_constants(*sb.step + 1, lr, b1, b2, eps, grad_scale);
    }
    __syncthreads();
    slice_pass<W, U>(ps, rank, sa, s_adam[0]);
    slice_pass<W, U>(ps, rank, sb, s_adam[1]);
    // ---- 5. replicas written: last CTA signals the peers and waits for theirs -----------------------------------------------
    __threadfence_system();
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(my_flags + FLAG_COUNT, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) my_flags[FLAG_COUNT] = 0;  // re-armed for the next step
    __threadfence_system();
    if (threadIdx.x < W && threadIdx.x != rank) st_release_sys(ps.flags[threadIdx.x] + flag_base + FLAG_DONE + rank, epoch);
    if (!wait_all(my_flags, FLAG_DONE, W, rank, epoch)) {
        if (threadIdx.x == 0) atomicExch(my_flags + FLAG_COUNT + 1, 2);
    }
}

__global__ void k_tick_step2(int* a, int* b) {
    *a += 1;
    if (b) *b += 1;
}

__global__ void k_tick_step(int* step) { *step += 1; }

typedef void (*exchange_fn)(const PeerSet, int, const SliceArgs, const SliceArgs, double, double, double, double, float, int, double, int);

extern "C" int nvo_exchange_flag_words(void) { return FLAG_WORDS; }

extern "C" int64_t nvo_exchange_slice(int64_t n, int32_t rank, int32_t world, int64_t* lo, int64_t* hi) {
    // slice of rank in floats: [lo, hi), cut on float4 boundaries; returns its length
    const int64_t n4 = (n + 3) / 4, chunk = (n4 + world - 1) / world;
    int64_t a = (int64_t)rank * chunk, b = a + chunk;
    if (a > n4) a = n4;
    if (b > n4) b = n4;
    if (lo) *lo = a * 4;
    if (hi) *hi = b * 4;
    return (b - a) * 4;
}

// Exchange + Adam of the flat ranges [offset, offset + n) (group A) and, if n_b > 0, [offset_b, offset_b + n_b) (group B) in ONE launch.
static int exchange_launch(void* stream, int64_t offset, int64_t n, float* m_a, float* v_a, int32_t* step_a, int64_t offset_b, int64_t n_b, float* m_b,
                           float* v_b, int32_t* step_b, int32_t phase, int32_t rank, int32_t world, const void* h_peer_params, const void* h_peer_grads,
                           const void* h_peer_flags, double lr, double beta1, double beta2, double eps, float grad_scale, int32_t ctas_per_sm,
                           double lr_final = 0.0, int max_steps = 1) {
    NVO_CHECK(n > 0 && (n & 3) == 0 && offset >= 0 && (offset & 3) == 0, "adam_exchange: range [%lld, +%lld) must be float4-aligned and non-empty",
              (long long)offset, (long long)n);
    NVO_CHECK(n_b >= 0 && (n_b & 3) == 0 && offset_b >= 0 && (offset_b & 3) == 0, "adam_exchange: second range [%lld, +%lld) must be float4-aligned",
              (long long)offset_b, (long long)n_b);
    NVO_CHECK(phase >= 0 && phase < NVO_MAX_PHASES, "adam_exchange: phase %d out of range [0,%d)", phase, NVO_MAX_PHASES);
    NVO_CHECK(world >= 1 && world <= NVO_MAX_PEERS && rank >= 0 && rank < world, "adam_exchange: bad rank/world %d/%d", rank, world);
    NVO_CHECK(h_peer_params && h_peer_grads && h_peer_flags && m_a && v_a && step_a, "adam_exchange: null pointer");
    NVO_CHECK(n_b == 0 || (m_b && v_b && step_b), "adam_exchange: null pointer (second group)");
    PeerSet ps;
    for (int k = 0; k < NVO_MAX_PEERS; ++k) {
        ps.params[k] = k < world ? ((float* const*)h_peer_params)[k] : nullptr;
        ps.grads[k] = k < world ? ((const float* const*)h_peer_grads)[k] : nullptr;
        ps.flags[k] = k < world ? ((int* const*)h_peer_flags)[k] : nullptr;
        if (k < world) NVO_CHECK(ps.params[k] && ps.grads[k] && ps.flags[k], "adam_exchange: null peer pointer for rank %d", k);
    }
    int64_t lo, hi, lob = 0, hib = 0;
    nvo_exchange_slice(n, rank, world, &lo, &hi);
    if (n_b > 0) nvo_exchange_slice(n_b, rank, world, &lob, &hib);
    const SliceArgs sa = {(offset + lo) / 4, (offset + hi) / 4, m_a, v_a, step_a};
    const SliceArgs sb = {(offset_b + lob) / 4, (offset_b + hib) / 4, m_b, v_b, n_b > 0 ? step_b : step_a};
    exchange_fn fn = nullptr;
    switch (world) {
        case 1: fn = k_exchange_adam<1, 4>; break;
        case 2: fn = k_exchange_adam<2, 4>; break;
        case 4: fn = k_exchange_adam<4, 2>; break;
        case 8: fn = k_exchange_adam<8, 1>; break;
        default: NVO_CHECK(false, "adam_exchange: world size %d unsupported (1, 2, 4 or 8 GPUs of one NVSwitch domain)", world);
    }
    cudaStream_t st = (cudaStream_t)stream;
    // persistent grid: as many 256-thread CTAs as are co-resident (measured at 2 GPUs: 151 us for the whole flat buffer against 175 us
    // with two CTAs per SM; at 8 GPUs the kernel is NVLink-bound and ONE CTA per SM is as fast).  ctas_per_sm > 0 (or
    // NVO_EXCHANGE_CTAS_PER_SM=k) caps it, for launches that run next to other kernels and must leave them registers.
    const int64_t items = (hi - lo) / 4 + (hib - lob) / 4;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, 0);
    per_sm = max(per_sm, 1);
    if (ctas_per_sm > 0) per_sm = min(per_sm, (int)ctas_per_sm);
    static const int cap = nvo_env_int("NVO_EXCHANGE_CTAS_PER_SM", 0);  // read once per process
    if (cap > 0) per_sm = max(1, min(per_sm, cap));
    const unsigned int grid = (unsigned int)max((int64_t)1, min((int64_t)nvo_sm_count() * per_sm, (items + 255) / 256));
    // An SM's L1 / shared-memory split is fixed while any CTA is resident.  This kernel uses no shared memory, so by default it would
    // configure its SMs with the smallest carve-out and the proposal backward (40 KB of dynamic shared memory per CTA) could not become
    // co-resident until the exchange had drained (observed: profiles/r01_timeline_n2_s9_early_no_carveout.csv).  Ask for the largest
    // carve-out instead.
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    fn<<<grid, 256, 0, st>>>(ps, rank, sa, sb, lr, beta1, beta2, eps, grad_scale, phase * FLAG_PHASE_STRIDE, lr_final, max_steps);
    NVO_CUDA_LAUNCH_CHECK("adam_exchange");
    k_tick_step2<<<1, 1, 0, st>>>(step_a, n_b > 0 ? step_b : nullptr);
    NVO_CUDA_LAUNCH_CHECK("adam_exchange(tick)");
    return 0;
}

extern "C" int nvo_adam_exchange_group(void* stream, int64_t offset, int64_t n, int32_t phase, int32_t rank, int32_t world, const void* h_peer_params,
                                       const void* h_peer_grads, const void* h_peer_flags, float* exp_avg_slice, float* exp_avg_sq_slice, int32_t* step,
                                       double lr, double beta1, double beta2, double eps, float grad_scale, int32_t ctas_per_sm) {
    return exchange_launch(stream, offset, n, exp_avg_slice, exp_avg_sq_slice, step, 0, 0, nullptr, nullptr, nullptr, phase, rank, world, h_peer_params,
                           h_peer_grads, h_peer_flags, lr, beta1, beta2, eps, grad_scale, ctas_per_sm);
}

// one group under ExponentialDecayScheduler (lr_init -> lr_final over max_steps, evaluated on the device from the group's step counter)
extern "C" int nvo_adam_exchange_group_decay(void* stream, int64_t offset, int64_t n, int32_t phase, int32_t rank, int32_t world,
                                             const void* h_peer_params, const void* h_peer_grads, const void* h_peer_flags, float* exp_avg_slice,
                                             float* exp_avg_sq_slice, int32_t* step, double lr_init, double lr_final, int32_t max_steps, double beta1,
                                             double beta2, double eps, float grad_scale, int32_t ctas_per_sm) {
    NVO_CHECK(lr_init > 0.0 && lr_final > 0.0 && max_steps >= 1, "adam_exchange_group_decay: bad schedule");
    return exchange_launch(stream, offset, n, exp_avg_slice, exp_avg_sq_slice, step, 0, 0, nullptr, nullptr, nullptr, phase, rank, world, h_peer_params,
                           h_peer_grads, h_peer_flags, lr_init, beta1, beta2, eps, grad_scale, ctas_per_sm, lr_final, max_steps);
}

extern "C" int nvo_adam_exchange_groups2(void* stream, int64_t offset_a, int64_t n_a, float* exp_avg_a, float* exp_avg_sq_a, int32_t* step_a,
                                         int64_t offset_b, int64_t n_b, float* exp_avg_b, float* exp_avg_sq_b, int32_t* step_b, int32_t rank,
                                         int32_t world, const void* h_peer_params, const void* h_peer_grads, const void* h_peer_flags, double lr,
                                         double beta1, double beta2, double eps, float grad_scale) {
    NVO_CHECK(n_b > 0, "adam_exchange_groups2: the second group is empty (use nvo_adam_exchange_group)");
    return exchange_launch(stream, offset_a, n_a, exp_avg_a, exp_avg_sq_a, step_a, offset_b, n_b, exp_avg_b, exp_avg_sq_b, step_b, 0, rank, world,
                           h_peer_params, h_peer_grads, h_peer_flags, lr, beta1, beta2, eps, grad_scale, 0);
}

extern "C" int nvo_adam_exchange_step(void* stream, int64_t n, int32_t rank, int32_t world, const void* h_peer_params, const void* h_peer_grads,
                                      const void* h_peer_flags, float* exp_avg_slice, float* exp_avg_sq_slice, int32_t* step, double lr, double beta1,
                                      double beta2, double eps, float grad_scale) {
    return nvo_adam_exchange_group(stream, 0, n, 0, rank, world, h_peer_params, h_peer_grads, h_peer_flags, exp_avg_slice, exp_avg_sq_slice, step, lr,
                                   beta1, beta2, eps, grad_scale, 0);
}

// ---- peer-visible allocations (CUDA IPC) ----------------------------------------------------------------------------------
extern "C" int nvo_peer_alloc(int64_t bytes, void* h_ptr_out, void* h_handle64_out) {
    NVO_CHECK(bytes > 0 && h_ptr_out && h_handle64_out, "peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    NVO_CHECK(e == cudaSuccess, "peer_alloc: cudaMalloc(%lld): %s", (long long)bytes, cudaGetErrorString(e));
    e = cudaMemset(p, 0, (size_t)bytes);
    NVO_CHECK(e == cudaSuccess, "peer_alloc: cudaMemset: %s", cudaGetErrorString(e));
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        NVO_CHECK(false, "peer_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    memcpy(h_handle64_out, &h, 64);
    *(void**)h_ptr_out = p;
    return 0;
}

extern "C" int nvo_peer_open(const void* h_handle64, void* h_ptr_out) {
    NVO_CHECK(h_handle64 && h_ptr_out, "peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    NVO_CHECK(e == cudaSuccess, "peer_open: cudaIpcOpenMemHandle: %s (peer access over NVLink is required for the fused exchange)", cudaGetErrorString(e));
    *(void**)h_ptr_out = p;
    return 0;
}

extern "C" int nvo_peer_close(void* ptr) {
    if (!ptr) return 0;
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    NVO_CHECK(e == cudaSuccess, "peer_close: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int nvo_peer_free(void* ptr) {
    if (!ptr) return 0;
    cudaError_t e = cudaFree(ptr);
    NVO_CHECK(e == cudaSuccess, "peer_free: %s", cudaGetErrorString(e));
    return 0;
}
