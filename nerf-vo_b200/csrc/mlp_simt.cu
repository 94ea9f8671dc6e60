// Generic fp32 SIMT MLP (any widths <= 64): the exact-arithmetic path used for narrow networks
// (proposal 10->16->1) and as the fp32 reference kernel the tensor-core path (mlp_tc.cu) is validated
// against on the device.  Semantics: NS/field_components/mlp.py:160-179 (Linear+bias, ReLU between layers,
// optional output activation), parameters in torch layout.
//
// Tile = 128 samples per CTA iteration, 256 threads: thread (s = tid & 127, h = tid >> 7) owns sample s and the
// odd/even 8-wide output chunks.  Activations live in shared memory transposed ([feature][sample], stride 129)
// so per-sample reads are conflict-free and weight reads are warp-broadcast float4s.
#include "nvo_common.cuh"

#define TILE 128
#define STRIDE 129
#define MAXW 64
#define NTHREADS 256

struct MlpP {
    int n_layers, in_dim;
    int dims[NVO_MAX_LAYERS], acts[NVO_MAX_LAYERS];
    int w_off[NVO_MAX_LAYERS], b_off[NVO_MAX_LAYERS], saved_off[NVO_MAX_LAYERS];
    int n_params, saved_per_sample;
};

__device__ __forceinline__ float act_fwd(float v, int act) {
    switch (act) {
        case NVO_ACT_RELU: return fmaxf(v, 0.f);
        case NVO_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case NVO_ACT_TANH: return tanhf(v);
        case NVO_ACT_EXP:
        case NVO_ACT_TRUNC_EXP: return expf(v);
        default: return v;
    }
}
// derivative expressed through the activation OUTPUT a
__device__ __forceinline__ float act_bwd_from_out(float a, int act) {
    switch (act) {
        case NVO_ACT_RELU: return a > 0.f ? 1.f : 0.f;
        case NVO_ACT_SIGMOID: return a * (1.f - a);
        case NVO_ACT_TANH: return 1.f - a * a;
        case NVO_ACT_EXP: return a;
        case NVO_ACT_TRUNC_EXP: return fminf(fmaxf(a, 3.0590232050182579e-07f), 3269017.3724721107f);  // exp(clamp(x,-15,15))
        default: return 1.f;
    }
}

// smem layout (floats): bufA[MAXW*STRIDE] | bufB[MAXW*STRIDE] | W[MAXW*MAXW] | bias[MAXW]
__global__ void __launch_bounds__(NTHREADS) k_mlp_fwd(const __grid_constant__ MlpP p, int64_t n, const float* __restrict__ x,
                                                      const float* __restrict__ params, const float* __restrict__ row_mask, float* __restrict__ y,
                                                      float* __restrict__ saved) {
    extern __shared__ float sm[];
    float* bufA = sm;
    float* bufB = sm + MAXW * STRIDE;
    float* Wt = bufB + MAXW * STRIDE;  // transposed [in][out_pad]
    float* bias = Wt + MAXW * MAXW;
    const int tid = threadIdx.x, s = tid & (TILE - 1), h = tid >> 7;
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * TILE;
        const int rows = (int)min((int64_t)TILE, n - base);
        // stage x tile transposed
        for (int e = tid; e < TILE * p.in_dim; e += NTHREADS) {
            const int r = e / p.in_dim, c = e - r * p.in_dim;
            bufA[c * STRIDE + r] = r < rows ? __ldg(x + (base + r) * p.in_dim + c) : 0.f;
        }
        float* in = bufA;
        float* out = bufB;
        int in_dim = p.in_dim;
        for (int l = 0; l < p.n_layers; ++l) {
            const int od = p.dims[l], od_pad = (od + 7) & ~7;
            const float* W = params + p.w_off[l];
            const float* B = params + p.b_off[l];
            __syncthreads();  // previous layer's reads of Wt done; `in` fully written
            for (int e = tid; e < in_dim * od_pad; e += NTHREADS) {
                const int i = e / od_pad, o = e - i * od_pad;
                Wt[e] = o < od ? __ldg(W + o * in_dim + i) : 0.f;
            }
            if (tid < od_pad) bias[tid] = tid < od ? __ldg(B + tid) : 0.f;
            __syncthreads();
            const bool last = l == p.n_layers - 1;
            const int act = p.acts[l];
            for (int c = h; c * 8 < od_pad; c += 2) {
                const int o0 = c * 8;
                float acc[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[u] = bias[o0 + u];
                for (int i = 0; i < in_dim; ++i) {
                    const float a = in[i * STRIDE + s];
                    const float4 w0 = *reinterpret_cast<const float4*>(Wt + i * od_pad + o0);
                    const float4 w1 = *reinterpret_cast<const float4*>(Wt + i * od_pad + o0 + 4);
                    acc[0] = fmaf(w0.x, a, acc[0]);
                    acc[1] = fmaf(w0.y, a, acc[1]);
                    acc[2] = fmaf(w0.z, a, acc[2]);
                    acc[3] = fmaf(w0.w, a, acc[3]);
                    acc[4] = fmaf(w1.x, a, acc[4]);
                    acc[5] = fmaf(w1.y, a, acc[5]);
                    acc[6] = fmaf(w1.z, a, acc[6]);
                    acc[7] = fmaf(w1.w, a, acc[7]);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int o = o0 + u;
                    if (o < od) {
                        const float v = act_fwd(acc[u], act);
                        out[o * STRIDE + s] = v;
                        if (s < rows) {
                            if (last)
                                y[(base + s) * od + o] = row_mask ? v * __ldg(row_mask + base + s) : v;
                            else if (saved)
                                saved[(base + s) * p.saved_per_sample + p.saved_off[l] + o] = v;
                        }
                    }
                }
            }
            float* t = in;
            in = out;
            out = t;
            in_dim = od;
        }
        __syncthreads();
    }
}

// Backward: persistent CTAs accumulate dW/db for all layers in shared memory over their tiles, then flush with
// one atomicAdd per parameter per CTA.
// smem (floats): A[MAXW*STRIDE] | G0[MAXW*STRIDE] | G1[MAXW*STRIDE] | W[MAXW*MAXW] | dP[n_params]
__global__ void __launch_bounds__(NTHREADS) k_mlp_bwd(const __grid_constant__ MlpP p, int64_t n, const float* __restrict__ x,
                                                      const float* __restrict__ params, const float* __restrict__ saved,
                                                      const float* __restrict__ y, const float* __restrict__ row_mask, const float* __restrict__ dy,
                                                      float* __restrict__ dx, float* __restrict__ dparams) {
    extern __shared__ float sm[];
    float* A = sm;
    float* G0 = A + MAXW * STRIDE;
    float* G1 = G0 + MAXW * STRIDE;
    float* Ws = G1 + MAXW * STRIDE;  // row-major [out][in]
    float* dP = Ws + MAXW * MAXW;
    const int tid = threadIdx.x, s = tid & (TILE - 1), h = tid >> 7;
    if (dparams)
        for (int e = tid; e < p.n_params; e += NTHREADS) dP[e] = 0.f;
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    const int last = p.n_layers - 1;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * TILE;
        const int rows = (int)min((int64_t)TILE, n - base);
        __syncthreads();
        // G = dL/dz of the last layer
        const int od_last = p.dims[last];
        for (int e = tid; e < TILE * od_last; e += NTHREADS) {
            const int r = e / od_last, c = e - r * od_last;
            float g = 0.f;
            if (r < rows) {
                g = __ldg(dy + (base + r) * od_last + c);
                if (row_mask) g *= __ldg(row_mask + base + r);
                if (p.acts[last] != NVO_ACT_NONE) g *= act_bwd_from_out(__ldg(y + (base + r) * od_last + c), p.acts[last]);
            }
            G0[c * STRIDE + r] = g;
        }
        float* G = G0;
        float* Gn = G1;
        for (int l = last; l >= 0; --l) {
            const int od = p.dims[l];
            const int id = l == 0 ? p.in_dim : p.dims[l - 1];
            __syncthreads();  // G complete; previous users of A / Ws done
            // stage the layer's input activations (transposed) and weights
            if (l == 0) {
                for (int e = tid; e < TILE * id; e += NTHREADS) {
                    const int r = e / id, c = e - r * id;
                    A[c * STRIDE + r] = r < rows ? __ldg(x + (base + r) * id + c) : 0.f;
                }
            } else {
                for (int e = tid; e < TILE * id; e += NTHREADS) {
                    const int r = e / id, c = e - r * id;
                    A[c * STRIDE + r] = r < rows ? __ldg(saved + (base + r) * p.saved_per_sample + p.saved_off[l - 1] + c) : 0.f;
                }
            }
            const bool need_dgrad = l > 0 || dx != nullptr;
            if (need_dgrad)
                for (int e = tid; e < od * id; e += NTHREADS) Ws[e] = __ldg(params + p.w_off[l] + e);
            __syncthreads();
            if (dparams) {
                // wgrad: thread owns elements e = tid, tid+256, ... of dW[l]; lanes -> consecutive i, same o (mostly)
                float* dW = dP + p.w_off[l];
                for (int e = tid; e < od * id; e += NTHREADS) {
                    const int o = e / id, i = e - o * id;
                    const float* g = G + o * STRIDE;
                    const float* a = A + i * STRIDE;
                    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 8
                    for (int r = 0; r < TILE; r += 2) {
                        acc0 = fmaf(g[r], a[r], acc0);
                        acc1 = fmaf(g[r + 1], a[r + 1], acc1);
                    }
                    dW[e] += acc0 + acc1;
                }
                if (tid < od) {
                    const float* g = G + tid * STRIDE;
                    float acc = 0.f;
                    for (int r = 0; r < TILE; ++r) acc += g[r];
                    dP[p.b_off[l] + tid] += acc;
                }
            }
            if (need_dgrad) {
                // dA[i][s] = sum_o W[o][i] G[o][s], times act'(A) for hidden inputs
                const int id_pad = (id + 7) & ~7;
                for (int c = h; c * 8 < id_pad; c += 2) {
                    const int i0 = c * 8;
                    float acc[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc[u] = 0.f;
                    for (int o = 0; o < od; ++o) {
                        const float g = G[o * STRIDE + s];
                        const float* w = Ws + o * id + i0;
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (i0 + u < id) acc[u] = fmaf(w[u], g, acc[u]);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + u;
                        if (i < id) {
                            if (l > 0) {
                                Gn[i * STRIDE + s] = acc[u] * act_bwd_from_out(A[i * STRIDE + s], p.acts[l - 1]);
                            } else if (s < rows) {
                                dx[(base + s) * id + i] = acc[u];
                            }
                        }
                    }
                }
                float* t = G;
                G = Gn;
                Gn = t;
            }
        }
    }
    __syncthreads();
    if (dparams)
        for (int e = tid; e < p.n_params; e += NTHREADS) atomicAdd(dparams + e, dP[e]);
}

static int make_params(const nvo_mlp_desc* d, MlpP* p) {
    NVO_CHECK(d != nullptr, "mlp: null descriptor");
    NVO_CHECK(d->n_layers >= 1 && d->n_layers <= NVO_MAX_LAYERS, "mlp: n_layers=%d out of range [1,%d]", d->n_layers, NVO_MAX_LAYERS);
    NVO_CHECK(d->in_dim >= 1 && d->in_dim <= MAXW, "mlp: in_dim=%d out of range [1,%d]", d->in_dim, MAXW);
    p->n_layers = d->n_layers;
    p->in_dim = d->in_dim;
    int off = 0, soff = 0, in = d->in_dim;
    for (int l = 0; l < NVO_MAX_LAYERS; ++l) {
        if (l < d->n_layers) {
            NVO_CHECK(d->dims[l] >= 1 && d->dims[l] <= MAXW, "mlp: layer %d width %d out of range [1,%d]", l, d->dims[l], MAXW);
            NVO_CHECK(d->acts[l] >= NVO_ACT_NONE && d->acts[l] <= NVO_ACT_TRUNC_EXP, "mlp: layer %d has unknown activation %d", l, d->acts[l]);
            p->dims[l] = d->dims[l];
            p->acts[l] = d->acts[l];
            p->w_off[l] = off;
            off += d->dims[l] * in;
            p->b_off[l] = off;
            off += d->dims[l];
            p->saved_off[l] = soff;
            if (l < d->n_layers - 1) soff += d->dims[l];
            in = d->dims[l];
        } else {
            p->dims[l] = p->acts[l] = p->w_off[l] = p->b_off[l] = p->saved_off[l] = 0;
        }
    }
    p->n_params = off;
    p->saved_per_sample = soff;
    return 0;
}

extern "C" int64_t nvo_mlp_n_params(const nvo_mlp_desc* d) {
    MlpP p;
    if (make_params(d, &p)) return -1;
    return p.n_params;
}
extern "C" int64_t nvo_mlp_saved_per_sample(const nvo_mlp_desc* d) {
    MlpP p;
    if (make_params(d, &p)) return -1;
    return p.saved_per_sample;
}

extern "C" int nvo_mlp_forward(const nvo_mlp_desc* d, void* stream, int64_t n, const float* x, const float* params, const float* row_mask, float* y,
                               float* saved) {
    MlpP p;
    if (int e = make_params(d, &p)) return e;
    NVO_CHECK(n >= 0, "mlp_forward: negative batch");
    if (n == 0) return 0;
    NVO_CHECK(x && params && y, "mlp_forward: null pointer");
    const size_t smem = sizeof(float) * (2 * MAXW * STRIDE + MAXW * MAXW + MAXW);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_mlp_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        NVO_CHECK(e == cudaSuccess, "mlp_forward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const int64_t tiles = (n + TILE - 1) / TILE;
    const unsigned int grid = (unsigned int)min(tiles, (int64_t)nvo_sm_count() * 2);
    k_mlp_fwd<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(p, n, x, params, row_mask, y, saved);
    NVO_CUDA_LAUNCH_CHECK("mlp_forward");
    return 0;
}

extern "C" int nvo_mlp_backward(const nvo_mlp_desc* d, void* stream, int64_t n, const float* x, const float* params, const float* saved,
                                const float* y, const float* row_mask, const float* dy, float* dx, float* dparams) {
    MlpP p;
    if (int e = make_params(d, &p)) return e;
    NVO_CHECK(n >= 0, "mlp_backward: negative batch");
    if (n == 0) return 0;
    NVO_CHECK(x && params && dy, "mlp_backward: null pointer");
    NVO_CHECK(p.n_layers == 1 || saved, "mlp_backward: saved activations required for multi-layer networks");
    NVO_CHECK(p.acts[p.n_layers - 1] == NVO_ACT_NONE || y, "mlp_backward: y required for an output activation");
    const size_t smem = sizeof(float) * (3 * MAXW * STRIDE + MAXW * MAXW + (size_t)p.n_params);
    NVO_CHECK(smem <= 227 * 1024, "mlp_backward: network too large for the SIMT path (%zu B shared)", smem);
    cudaError_t e = cudaFuncSetAttribute(k_mlp_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    NVO_CHECK(e == cudaSuccess, "mlp_backward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    const int64_t tiles = (n + TILE - 1) / TILE;
    const unsigned int grid = (unsigned int)min(tiles, (int64_t)nvo_sm_count());
    k_mlp_bwd<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(p, n, x, params, saved, y, row_mask, dy, dx, dparams);
    NVO_CUDA_LAUNCH_CHECK("mlp_backward");
    return 0;
}
