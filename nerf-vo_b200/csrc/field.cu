// Element-wise field operators: scene contraction + normalisation, SH / frequency encodings, trunc_exp, 3-vector
// normalisation, and the fused input assembly of NerfactoField.get_outputs (density activation, SH of the view
// direction, position encoding, appearance-embedding gather, concatenation) with its backward.
#include "nvo_common.cuh"

// ---- contraction (spatial_distortions.py:67-69) + (x+2)/4 + selector (nerfacto_field.py:204-209) -----------------
__global__ void __launch_bounds__(256) k_contract(int64_t n, const float* __restrict__ pos, float* __restrict__ x, float* __restrict__ selector) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float p[3] = {__ldg(pos + 3 * t), __ldg(pos + 3 * t + 1), __ldg(pos + 3 * t + 2)};
    const float mag = fmaxf(fabsf(p[0]), fmaxf(fabsf(p[1]), fabsf(p[2])));
    bool sel = true;
    float q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float c = p[a];
        if (!(mag < 1.f)) c = __fmul_rn(__fsub_rn(2.f, __fdiv_rn(1.f, mag)), __fdiv_rn(p[a], mag));
        q[a] = __fmul_rn(__fadd_rn(c, 2.f), 0.25f);
        sel = sel && (q[a] > 0.f) && (q[a] < 1.f);
    }
    const float m = sel ? 1.f : 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) x[3 * t + a] = __fmul_rn(q[a], m);
    selector[t] = m;
}

// sample positions (rays.py:55: origins + directions * (starts + ends) / 2) and their contraction in one pass: the fused field forward needs
// both (world positions feed the predicted-normals encoding, contracted ones the hash grid); same arithmetic as k_sample_positions + k_contract
__global__ void __launch_bounds__(256) k_positions_contract(int64_t B, int S, const float* __restrict__ o, const float* __restrict__ d,
                                                            const float* __restrict__ starts, const float* __restrict__ ends, int64_t stride,
                                                            float* __restrict__ pos, float* __restrict__ x, float* __restrict__ selector) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * S) return;
    const int64_t r = t / S;
    const int k = (int)(t - r * S);
    const float se = __fadd_rn(__ldg(starts + r * stride + k), __ldg(ends + r * stride + k));
    float p[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        p[a] = __fadd_rn(__ldg(o + 3 * r + a), __fmul_rn(__fmul_rn(__ldg(d + 3 * r + a), se), 0.5f));
        pos[3 * t + a] = p[a];
    }
    const float mag = fmaxf(fabsf(p[0]), fmaxf(fabsf(p[1]), fabsf(p[2])));
    bool sel = true;
    float q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float c = p[a];
        if (!(mag < 1.f)) c = __fmul_rn(__fsub_rn(2.f, __fdiv_rn(1.f, mag)), __fdiv_rn(p[a], mag));
        q[a] = __fmul_rn(__fadd_rn(c, 2.f), 0.25f);
        sel = sel && (q[a] > 0.f) && (q[a] < 1.f);
    }
    const float m = sel ? 1.f : 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) x[3 * t + a] = __fmul_rn(q[a], m);
    selector[t] = m;
}

// ---- SH degree 4 (NS/utils/math.py:45-78), constants rounded to fp32 as torch does for python-float * tensor ----
__device__ __forceinline__ void sh16(float x, float y, float z, float* c) {
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

__global__ void __launch_bounds__(256) k_sh4(int64_t n, const float* __restrict__ d, float* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float c[16];
    sh16(__ldg(d + 3 * t), __ldg(d + 3 * t + 1), __ldg(d + 3 * t + 2), c);
#pragma unroll
    for (int k = 0; k < 16; k += 4) reinterpret_cast<float4*>(out + 16 * t)[k >> 2] = make_float4(c[k], c[k + 1], c[k + 2], c[k + 3]);
}

// ---- frequency encoding, torch path (encodings.py:170-176) ------------------------------------------------------
__device__ __forceinline__ float posenc_value(float xi, int k, bool cos_block) {
    // u = (2*pi*x) * 2^k ; sin(u) or sin(u + pi/2)
    const float u = __fmul_rn(__fmul_rn(6.283185307179586f, xi), (float)(1 << k));
    return sinf(cos_block ? __fadd_rn(u, 1.5707963267948966f) : u);
}

__global__ void __launch_bounds__(256) k_frequency(int64_t n, int in_dim, int n_freq, const float* __restrict__ x, float* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int half = in_dim * n_freq, od = 2 * half;
    if (t >= n * od) return;
    const int64_t r = t / od;
    int c = (int)(t - r * od);
    const bool cosb = c >= half;
    if (cosb) c -= half;
    const int i = c / n_freq, k = c - i * n_freq;
    out[t] = posenc_value(__ldg(x + r * in_dim + i), k, cosb);
}

// ---- trunc_exp -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_trunc_exp_fwd(int64_t n, const float* __restrict__ x, float* __restrict__ y) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) y[t] = expf(__ldg(x + t));
}
__global__ void __launch_bounds__(256) k_trunc_exp_bwd(int64_t n, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dx[t] = __ldg(dy + t) * expf(fminf(fmaxf(__ldg(x + t), -15.f), 15.f));
}

// ---- normalize3 --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_normalize3_fwd(int64_t n, const float* __restrict__ v, float scale, float eps, float* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float a = __ldg(v + 3 * t), b = __ldg(v + 3 * t + 1), c = __ldg(v + 3 * t + 2);
    const float nrm = fmaxf(sqrtf(a * a + b * b + c * c), eps);
    out[3 * t] = scale * (a / nrm);
    out[3 * t + 1] = scale * (b / nrm);
    out[3 * t + 2] = scale * (c / nrm);
}
__global__ void __launch_bounds__(256) k_normalize3_bwd(int64_t n, const float* __restrict__ v, const float* __restrict__ dout, float scale, float eps,
                                                        float* __restrict__ dv) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float a = __ldg(v + 3 * t), b = __ldg(v + 3 * t + 1), c = __ldg(v + 3 * t + 2);
    const float g0 = scale * __ldg(dout + 3 * t), g1 = scale * __ldg(dout + 3 * t + 1), g2 = scale * __ldg(dout + 3 * t + 2);
    const float r = sqrtf(a * a + b * b + c * c);
    if (r > eps) {
        const float dot = (a * g0 + b * g1 + c * g2) / (r * r * r);
        dv[3 * t] = g0 / r - a * dot;
        dv[3 * t + 1] = g1 / r - b * dot;
        dv[3 * t + 2] = g2 / r - c * dot;
    } else {  // clamped denominator is a constant
        dv[3 * t] = g0 / eps;
        dv[3 * t + 1] = g1 / eps;
        dv[3 * t + 2] = g2 / eps;
    }
}

// ---- fused head-input assembly --------------------------------------------------------------------------------
#define GEO 15
#define APP 32
#define HEAD_IN 63
#define PN_IN 27

// head_in [n,63] = [SH16((dir+1)/2) | h[1:16] | appearance32], pn_in [n,27] = [posenc12(pos) | h[1:16]] for one sample
__device__ __forceinline__ void assemble_rows(int64_t t, int64_t r, const float* __restrict__ h, const float* __restrict__ dirs,
                                              const float* __restrict__ pos, const int64_t* __restrict__ cam_idx, const float* __restrict__ embedding,
                                              float* hv, float* head, float* pn, bool want_pn) {
#pragma unroll
    for (int k = 0; k < 16; k += 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(h + 16 * t) + (k >> 2));
        hv[k] = q.x, hv[k + 1] = q.y, hv[k + 2] = q.z, hv[k + 3] = q.w;
    }
    // get_normalized_directions (base_field.py:142): (d + 1) / 2
    sh16(__fmul_rn(__fadd_rn(__ldg(dirs + 3 * r), 1.f), 0.5f), __fmul_rn(__fadd_rn(__ldg(dirs + 3 * r + 1), 1.f), 0.5f),
         __fmul_rn(__fadd_rn(__ldg(dirs + 3 * r + 2), 1.f), 0.5f), head);
#pragma unroll
    for (int k = 0; k < GEO; ++k) head[16 + k] = hv[1 + k];
    const float* e = cam_idx ? embedding + APP * __ldg(cam_idx + r) : embedding;
#pragma unroll
    for (int k = 0; k < APP; k += 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(e + k));
        head[16 + GEO + k] = q.x, head[16 + GEO + k + 1] = q.y, head[16 + GEO + k + 2] = q.z, head[16 + GEO + k + 3] = q.w;
    }
    head[HEAD_IN] = 0.f;
    if (want_pn) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float xi = __ldg(pos + 3 * t + i);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                pn[i * 2 + k] = posenc_value(xi, k, false);
                pn[6 + i * 2 + k] = posenc_value(xi, k, true);
            }
        }
#pragma unroll
        for (int k = 0; k < GEO; ++k) pn[12 + k] = hv[1 + k];
#pragma unroll
        for (int k = PN_IN; k < 32; ++k) pn[k] = 0.f;
    }
}

// fp32 outputs: head_in [n,63], pn_in [n,27] row-major (the reference's widths)
__global__ void __launch_bounds__(256)
    k_assemble_fwd(int64_t B, int S, const float* __restrict__ h, const float* __restrict__ selector, const float* __restrict__ dirs,
                   const float* __restrict__ pos, const int64_t* __restrict__ cam_idx, const float* __restrict__ embedding, float* __restrict__ density,
                   float* __restrict__ head_in, float* __restrict__ pn_in) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * S) return;
    const int64_t r = t / S;
    float hv[16], head[64], pn[32];
    assemble_rows(t, r, h, dirs, pos, cam_idx, embedding, hv, head, pn, pn_in != nullptr);
    if (density) density[t] = __fmul_rn(expf(hv[0]), __ldg(selector + t));
#pragma unroll
    for (int k = 0; k < HEAD_IN; ++k) head_in[(int64_t)HEAD_IN * t + k] = head[k];
    if (pn_in) {
#pragma unroll
        for (int k = 0; k < PN_IN; ++k) pn_in[(int64_t)PN_IN * t + k] = pn[k];
    }
}

__device__ __forceinline__ uint4 pack8h(const float* v) {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]), c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    u.z = *reinterpret_cast<uint32_t*>(&c);
    u.w = *reinterpret_cast<uint32_t*>(&d);
    return u;
}

// fp16 outputs in the tensor-core MLP's operand layout (TMH tiles of 128 rows, zero padded to 64 / 32 columns and to whole
// tiles): thread == row, every 8-feature chunk is one 16-byte store, a warp writes 512 contiguous bytes per chunk.
__global__ void __launch_bounds__(128)
    k_assemble_fwd_tmh(int64_t B, int S, const float* __restrict__ h, const float* __restrict__ selector, const float* __restrict__ dirs,
                       const float* __restrict__ pos, const int64_t* __restrict__ cam_idx, const float* __restrict__ embedding, float* __restrict__ density,
                       uint4* __restrict__ head_in, uint4* __restrict__ pn_in) {
    const int64_t tile = blockIdx.x;
    const int rr = threadIdx.x;
    const int64_t t = tile * 128 + rr;
    float hv[16], head[64], pn[32];
    if (t < B * S) {
        assemble_rows(t, t / S, h, dirs, pos, cam_idx, embedding, hv, head, pn, pn_in != nullptr);
        if (density) density[t] = __fmul_rn(expf(hv[0]), __ldg(selector + t));
    } else {
#pragma unroll
        for (int k = 0; k < 64; ++k) head[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) pn[k] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) head_in[(tile * 8 + c) * 128 + rr] = pack8h(head + 8 * c);
    if (pn_in) {
#pragma unroll
        for (int c = 0; c < 4; ++c) pn_in[(tile * 4 + c) * 128 + rr] = pack8h(pn + 8 * c);
    }
}

__global__ void __launch_bounds__(256)
    k_assemble_bwd(int64_t B, int S, const float* __restrict__ h, const float* __restrict__ selector, const int64_t* __restrict__ cam_idx,
                   const float* __restrict__ ddensity, const float* __restrict__ dhead_in, const float* __restrict__ dpn_in, int tmf,
                   float* __restrict__ dh, float* __restrict__ dembedding, int* __restrict__ dh_absmax_bits) {
    // element (sample t, column c) of a [n, K] gradient: row-major, or TMF [tile][K][128] (mlp_tc.cu's dx layout)
    auto at = [tmf](const float* g, int K, int64_t t, int c) { return tmf ? __ldg(g + (((t >> 7) * K + c) << 7) + (t & 127)) : __ldg(g + K * t + c); };
    // one warp per ray: lanes stride over the ray's samples (every global access of the warp is a contiguous run in the tile-major
    // layout); each lane sums the 32 appearance-gradient channels of ITS samples in registers, a butterfly transpose-reduce (31 shuffles)
    // then leaves channel k in lane k: ONE atomicAdd per (ray, channel) instead of S.
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= B) return;
    const bool want_emb = dembedding && cam_idx;
    float e[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) e[k] = 0.f;
    float amax = 0.f;  // max |dh| of this lane's samples: the consumer (tensor-core MLP backward) derives its gradient scale from it
    for (int s0 = 0; s0 < S; s0 += 32) {
        const int s = s0 + lane;
        if (s < S) {
            const int64_t t = r * S + s;
            float g[16];
            // trunc_exp backward (activations.py:38-41) with the selector product
            const float h0 = __ldg(h + 16 * t);
            g[0] = ddensity ? __ldg(ddensity + t) * __ldg(selector + t) * expf(fminf(fmaxf(h0, -15.f), 15.f)) : 0.f;
#pragma unroll
            for (int k = 0; k < GEO; ++k) {
                float v = at(dhead_in, HEAD_IN, t, 16 + k);
                if (dpn_in) v += at(dpn_in, PN_IN, t, 12 + k);
                g[1 + k] = v;
            }
#pragma unroll
            for (int k = 0; k < 16; k += 4) reinterpret_cast<float4*>(dh + 16 * t)[k >> 2] = make_float4(g[k], g[k + 1], g[k + 2], g[k + 3]);
            if (dh_absmax_bits) {
#pragma unroll
                for (int k = 0; k < 16; ++k) amax = fmaxf(amax, fabsf(g[k]));
            }
            if (want_emb) {
#pragma unroll
                for (int k = 0; k < 32; ++k) e[k] += at(dhead_in, HEAD_IN, t, 16 + GEO + k);
            }
        }
    }
    if (dh_absmax_bits) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        // non-negative floats order like their bit patterns (k_absmax_scale's convention); NaN / Inf are left to the consumer's guard
        if (lane == 0 && amax > 0.f) atomicMax(dh_absmax_bits, __float_as_int(amax));
    }
    if (want_emb) {
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
        atomicAdd(dembedding + APP * __ldg(cam_idx + r) + lane, e[0]);
    }
}

// ---- C ABI ----------------------------------------------------------------------------------------------------
extern "C" int nvo_contract_forward(void* stream, int64_t n, const float* pos, float* x, float* selector) {
    NVO_CHECK(n >= 0, "contract_forward: negative size");
    if (n == 0) return 0;
    NVO_CHECK(pos && x && selector, "contract_forward: null pointer");
    k_contract<<<nvo_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(n, pos, x, selector);
    NVO_CUDA_LAUNCH_CHECK("contract_forward");
    return 0;
}
extern "C" int nvo_sample_positions_contract(void* stream, int64_t B, int32_t S, const float* origins, const float* directions, const float* starts,
                                             const float* ends, int64_t stride, float* positions, float* x, float* selector) {
    NVO_CHECK(B >= 0 && S >= 1, "sample_positions_contract: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(origins && directions && starts && ends && positions && x && selector, "sample_positions_contract: null pointer");
    k_positions_contract<<<nvo_blocks(B * S, 256), 256, 0, (cudaStream_t)stream>>>(B, S, origins, directions, starts, ends, stride, positions, x, selector);
    NVO_CUDA_LAUNCH_CHECK("sample_positions_contract");
    return 0;
}
extern "C" int nvo_sh4_forward(void* stream, int64_t n, const float* d, float* out) {
    NVO_CHECK(n >= 0, "sh4_forward: negative size");
    if (n == 0) return 0;
    NVO_CHECK(d && out, "sh4_forward: null pointer");
    k_sh4<<<nvo_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(n, d, out);
    NVO_CUDA_LAUNCH_CHECK("sh4_forward");
    return 0;
}
extern "C" int nvo_frequency_forward(void* stream, int64_t n, int32_t in_dim, int32_t n_freq, const float* x, float* out) {
    NVO_CHECK(n >= 0 && in_dim >= 1 && n_freq >= 1 && n_freq <= 24, "frequency_forward: bad shape");
    if (n == 0) return 0;
    NVO_CHECK(x && out, "frequency_forward: null pointer");
    k_frequency<<<nvo_blocks(n * in_dim * n_freq * 2, 256), 256, 0, (cudaStream_t)stream>>>(n, in_dim, n_freq, x, out);
    NVO_CUDA_LAUNCH_CHECK("frequency_forward");
    return 0;
}
extern "C" int nvo_trunc_exp_forward(void* stream, int64_t n, const float* x, float* y) {
    NVO_CHECK(n >= 0, "trunc_exp_forward: negative size");
    if (n == 0) return 0;
    NVO_CHECK(x && y, "trunc_exp_forward: null pointer");
    k_trunc_exp_fwd<<<nvo_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(n, x, y);
    NVO_CUDA_LAUNCH_CHECK("trunc_exp_forward");
    return 0;
}
extern "C" int nvo_trunc_exp_backward(void* stream, int64_t n, const float* x, const float* dy, float* dx) {
    NVO_CHECK(n >= 0, "trunc_exp_backward: negative size");
    if (n == 0) return 0;
    NVO_CHECK(x && dy && dx, "trunc_exp_backward: null pointer");
    k_trunc_exp_bwd<<<nvo_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(n, x, dy, dx);
    NVO_CUDA_LAUNCH_CHECK("trunc_exp_backward");
    return 0;
}
extern "C" int nvo_normalize3_forward(void* stream, int64_t n, const float* v, float scale, float eps, float* out) {
    NVO_CHECK(n >= 0, "normalize3_forward: negative size");
    if (n == 0) return 0;
    NVO_CHECK(v && out, "normalize3_forward: null pointer");
    k_normalize3_fwd<<<nvo_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(n, v, scale, eps, out);
    NVO_CUDA_LAUNCH_CHECK("normalize3_forward");
    return 0;
}
extern "C" int nvo_normalize3_backward(void* stream, int64_t n, const float* v, const float* dout, float scale, float eps, float* dv) {
    NVO_CHECK(n >= 0, "normalize3_backward: negative size");
    if (n == 0) return 0;
    NVO_CHECK(v && dout && dv, "normalize3_backward: null pointer");
    k_normalize3_bwd<<<nvo_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(n, v, dout, scale, eps, dv);
    NVO_CUDA_LAUNCH_CHECK("normalize3_backward");
    return 0;
}
extern "C" int nvo_field_assemble_forward(void* stream, int64_t B, int32_t S, const float* h, const float* selector, const float* directions,
                                          const float* pos, const int64_t* cam_idx, const float* embedding, int32_t f16_padded, float* density,
                                          void* head_in, void* pn_in) {
    NVO_CHECK(B >= 0 && S >= 1, "field_assemble_forward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(h && directions && embedding && head_in, "field_assemble_forward: null pointer");
    NVO_CHECK(!density || selector, "field_assemble_forward: selector required for density");
    NVO_CHECK(!pn_in || pos, "field_assemble_forward: positions required for pn_in");
    if (f16_padded)
        k_assemble_fwd_tmh<<<(unsigned int)((B * S + 127) / 128), 128, 0, (cudaStream_t)stream>>>(B, S, h, selector, directions, pos, cam_idx, embedding, density,
                                                                                                 (uint4*)head_in, (uint4*)pn_in);
    else
        k_assemble_fwd<<<nvo_blocks(B * S, 256), 256, 0, (cudaStream_t)stream>>>(B, S, h, selector, directions, pos, cam_idx, embedding, density,
                                                                                  (float*)head_in, (float*)pn_in);
    NVO_CUDA_LAUNCH_CHECK("field_assemble_forward");
    return 0;
}
extern "C" int nvo_field_assemble_backward(void* stream, int64_t B, int32_t S, const float* h, const float* selector, const int64_t* cam_idx,
                                           const float* ddensity, const float* dhead_in, const float* dpn_in, int32_t tmf, float* dh, float* dembedding,
                                           float* dh_absmax) {
    NVO_CHECK(B >= 0 && S >= 1, "field_assemble_backward: bad shape");
    if (B == 0) return 0;
    NVO_CHECK(h && dhead_in && dh, "field_assemble_backward: null pointer");
    NVO_CHECK(!ddensity || selector, "field_assemble_backward: selector required for ddensity");
    k_assemble_bwd<<<nvo_blocks(B * 32, 256), 256, 0, (cudaStream_t)stream>>>(B, S, h, selector, cam_idx, ddensity, dhead_in, dpn_in, tmf, dh, dembedding,
                                                                              reinterpret_cast<int*>(dh_absmax));
    NVO_CUDA_LAUNCH_CHECK("field_assemble_backward");
    return 0;
}
