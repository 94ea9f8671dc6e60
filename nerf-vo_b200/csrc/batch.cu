// Step prologue: pixel sampling, pixel gather, pinhole ray generation and camera-pose correction fused into one
// launch (one thread per ray), plus the exp-map / pose-gradient kernels of the camera optimizer.
//   reference: nerf_vo/mapping/nerfstudio_utils.py:133-155,295-300; NS/data/pixel_samplers.py:103-106,170-219;
//   NS/model_components/ray_generators.py:40-57; NS/cameras/cameras.py:596-654,780-785,865-912;
//   NS/cameras/camera_optimizers.py:108-147; NS/cameras/lie_groups.py:25-120.
#include "nvo_common.cuh"

namespace {

// ---- forward-mode dual number: one templated exp map serves the forward (float) and the pose gradient (Dual) ---------------------
struct Dual {
    float v, d;
};
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual operator-(Dual a) { return {-a.v, -a.d}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
    const float q = a.v / b.v;
    return {q, (a.d - q * b.d) / b.v};
}
__device__ __forceinline__ Dual nsin(Dual a) { return {sinf(a.v), cosf(a.v) * a.d}; }
__device__ __forceinline__ Dual ncos(Dual a) { return {cosf(a.v), -sinf(a.v) * a.d}; }
__device__ __forceinline__ Dual nsqrt(Dual a) {
    const float r = sqrtf(a.v);
    return {r, 0.5f * a.d / r};
}
__device__ __forceinline__ float nsin(float a) { return sinf(a); }
__device__ __forceinline__ float ncos(float a) { return cosf(a); }
__device__ __forceinline__ float nsqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ float val(float a) { return a; }
__device__ __forceinline__ float val(Dual a) { return a.v; }
template <typename T>
__device__ __forceinline__ T cst(float c);
template <>
__device__ __forceinline__ float cst<float>(float c) { return c; }
template <>
__device__ __forceinline__ Dual cst<Dual>(float c) { return {c, 0.f}; }

// M[3][4] = exp map of tangent t[6] = (translation, rotation)
template <typename T>
__device__ __forceinline__ void exp_map(int mode, const T* t, T (*M)[4]) {
    const T w0 = t[3], w1 = t[4], w2 = t[5];
    const T n2 = w0 * w0 + w1 * w1 + w2 * w2;
    T A, Bc, C = cst<T>(0.f);
    if (mode == NVO_POSE_SO3XR3) {
        // lie_groups.py:37-41: angle = sqrt(clamp(|w|^2, 1e-4)); below the clamp the factors are constants
        const T ang = val(n2) < 1e-4f ? cst<T>(0.01f) : nsqrt(n2);
        const T inv = cst<T>(1.f) / ang;
        A = inv * nsin(ang);
        Bc = inv * inv * (cst<T>(1.f) - ncos(ang));
    } else {
        if (val(n2) < 1e-8f) {  // series of sin(th)/th, (1-cos th)/th^2, (th-sin th)/th^3
            A = cst<T>(1.f) - n2 * cst<T>(1.f / 6.f);
            Bc = cst<T>(0.5f) - n2 * cst<T>(1.f / 24.f);
            C = cst<T>(1.f / 6.f) - n2 * cst<T>(1.f / 120.f);
        } else {
            const T th = nsqrt(n2);
            A = nsin(th) / th;
            // 1 - cos(th) = 2 sin^2(th/2): no cancellation for small angles
            const T sh = nsin(th * cst<T>(0.5f));
            Bc = cst<T>(2.f) * sh * sh / n2;
            C = val(n2) < 1e-2f ? cst<T>(1.f / 6.f) - n2 * (cst<T>(1.f / 120.f) - n2 * cst<T>(1.f / 5040.f)) : (th - nsin(th)) / (n2 * th);
        }
    }
    // K = [w]x, K^2 = w w^T - |w|^2 I
    const T z = cst<T>(0.f), one = cst<T>(1.f);
    const T Kx[3][3] = {{z, -w2, w1}, {w2, z, -w0}, {-w1, w0, z}};
    const T w[3] = {w0, w1, w2};
    T K2[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) K2[i][j] = w[i] * w[j] - (i == j ? n2 : z);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) M[i][j] = A * Kx[i][j] + Bc * K2[i][j] + (i == j ? one : z);
    if (mode == NVO_POSE_SO3XR3) {
#pragma unroll
        for (int i = 0; i < 3; ++i) M[i][3] = t[i];
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            T acc = z;
#pragma unroll
            for (int j = 0; j < 3; ++j) acc = acc + (Bc * Kx[i][j] + C * K2[i][j] + (i == j ? one : z)) * t[j];
            M[i][3] = acc;
        }
    }
}

// ---- pinhole ray through the centre of pixel (py,px) of frame c (cameras.py:596-633,654,780-785,865-892) -------------------------
__device__ __forceinline__ void pinhole_ray(const float* __restrict__ intr, const float* __restrict__ ext, int64_t c, int64_t py, int64_t px,
                                            float* o, float* d, float* dnorm, float* area) {
    const float4 in = __ldg(reinterpret_cast<const float4*>(intr) + c);  // fx fy cx cy
    const float4* E = reinterpret_cast<const float4*>(ext) + 4 * c;
    const float4 r0 = __ldg(E), r1 = __ldg(E + 1), r2 = __ldg(E + 2);
    const float x = (float)px + 0.5f, y = (float)py + 0.5f;
    const float xc = __fsub_rn(x, in.z), yc = __fsub_rn(y, in.w);
    // coord, coord_x_offset, coord_y_offset; y negated (OpenCV -> OpenGL), z = -1
    const float cx0 = __fdiv_rn(xc, in.x), cy0 = -__fdiv_rn(yc, in.y);
    const float cx1 = __fdiv_rn(__fadd_rn(xc, 1.f), in.x), cy2 = -__fdiv_rn(__fadd_rn(yc, 1.f), in.y);
    float dir[3][3];
    const float cam[3][2] = {{cx0, cy0}, {cx1, cy0}, {cx0, cy2}};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float a = cam[k][0], b = cam[k][1];
        float v0 = __fadd_rn(__fadd_rn(__fmul_rn(a, r0.x), __fmul_rn(b, r0.y)), -r0.z);
        float v1 = __fadd_rn(__fadd_rn(__fmul_rn(a, r1.x), __fmul_rn(b, r1.y)), -r1.z);
        float v2 = __fadd_rn(__fadd_rn(__fmul_rn(a, r2.x), __fmul_rn(b, r2.y)), -r2.z);
        const float nrm = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v0, v0), __fmul_rn(v1, v1)), __fmul_rn(v2, v2))), 1e-8f);
        dir[k][0] = __fdiv_rn(v0, nrm);
        dir[k][1] = __fdiv_rn(v1, nrm);
        dir[k][2] = __fdiv_rn(v2, nrm);
        if (k == 0) *dnorm = nrm;
    }
    float dd[2];
#pragma unroll
    for (int k = 1; k < 3; ++k) {
        const float e0 = __fsub_rn(dir[0][0], dir[k][0]), e1 = __fsub_rn(dir[0][1], dir[k][1]), e2 = __fsub_rn(dir[0][2], dir[k][2]);
        dd[k - 1] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1)), __fmul_rn(e2, e2)));
    }
    *area = __fmul_rn(dd[0], dd[1]);
    o[0] = r0.w, o[1] = r1.w, o[2] = r2.w;
    d[0] = dir[0][0], d[1] = dir[0][1], d[2] = dir[0][2];
}

__device__ __forceinline__ void store3(float* p, int64_t i, const float* v) {
    p[3 * i] = v[0], p[3 * i + 1] = v[1], p[3 * i + 2] = v[2];
}

__global__ void __launch_bounds__(128) k_batch_prologue(int64_t B, int K, const int32_t* __restrict__ K_dev, int H, int W, const float* __restrict__ u,
                                                        const float* __restrict__ intr,
                                                        const float* __restrict__ ext, const float* __restrict__ color, const float* __restrict__ depthf,
                                                        const float* __restrict__ normalf, const float* __restrict__ pose, int mode,
                                                        int64_t* __restrict__ indices, int64_t* __restrict__ cam_idx, float* __restrict__ origins,
                                                        float* __restrict__ directions, float* __restrict__ dnorm, float* __restrict__ area,
                                                        float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ normal,
                                                        float* __restrict__ draw) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    // the number of active keyframes as a device scalar: a captured CUDA graph keeps sampling the frames the mapping thread adds
    // (clamped to the host-validated capacity K)
    if (K_dev) K = max(1, min(K, __ldg(K_dev)));
    // pixel_samplers.py:103-106: (rand * [K,H,W]).long()
    // (u < 1 can still round u*K up to K in fp32: the reference then indexes out of bounds; clamped here for memory safety)
    const int64_t c = min((int64_t)__fmul_rn(__ldg(u + 3 * i), (float)K), (int64_t)K - 1);
    const int64_t py = min((int64_t)__fmul_rn(__ldg(u + 3 * i + 1), (float)H), (int64_t)H - 1);
    const int64_t px = min((int64_t)__fmul_rn(__ldg(u + 3 * i + 2), (float)W), (int64_t)W - 1);
    indices[3 * i] = c, indices[3 * i + 1] = py, indices[3 * i + 2] = px;
    cam_idx[i] = c;
    const int64_t pix = (c * H + py) * W + px;
    // the three gathers are issued before the ray arithmetic so their latency overlaps it
    const float cr = __ldg(color + 3 * pix), cg = __ldg(color + 3 * pix + 1), cb = __ldg(color + 3 * pix + 2);
    const float dp = __ldg(depthf + pix);
    float n[3] = {0.f, 0.f, 0.f};
    if (normalf) n[0] = __ldg(normalf + 3 * pix), n[1] = __ldg(normalf + 3 * pix + 1), n[2] = __ldg(normalf + 3 * pix + 2);
    float o[3], d[3], nrm, ar;
    pinhole_ray(intr, ext, c, py, px, o, d, &nrm, &ar);
    if (draw) store3(draw, i, d);
    if (mode != NVO_POSE_OFF) {
        float t[6], M[3][4];
#pragma unroll
        for (int k = 0; k < 6; ++k) t[k] = __ldg(pose + 6 * c + k);
        exp_map<float>(mode, t, M);
        float d2[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            o[a] += M[a][3];
            d2[a] = M[a][0] * d[0] + M[a][1] * d[1] + M[a][2] * d[2];
        }
        d[0] = d2[0], d[1] = d2[1], d[2] = d2[2];
    }
    store3(origins, i, o);
    store3(directions, i, d);
    dnorm[i] = nrm;
    area[i] = ar;
    rgb[3 * i] = cr, rgb[3 * i + 1] = cg, rgb[3 * i + 2] = cb;
    depth[i] = dp;
    if (normalf) {
        // nerfstudio_utils.py:142-150: (solve(R, n) + 1) / 2 with R = extrinsics[c,:3,:3] (not assumed orthonormal): adjugate / determinant
        const float4* E = reinterpret_cast<const float4*>(ext) + 4 * c;
        const float4 a = __ldg(E), b = __ldg(E + 1), g = __ldg(E + 2);
        const float c00 = b.y * g.z - b.z * g.y, c01 = b.z * g.x - b.x * g.z, c02 = b.x * g.y - b.y * g.x;
        const float det = a.x * c00 + a.y * c01 + a.z * c02;
        const float inv = 1.f / det;
        const float s0 = (c00 * n[0] + (a.z * g.y - a.y * g.z) * n[1] + (a.y * b.z - a.z * b.y) * n[2]) * inv;
        const float s1 = (c01 * n[0] + (a.x * g.z - a.z * g.x) * n[1] + (a.z * b.x - a.x * b.z) * n[2]) * inv;
        const float s2 = (c02 * n[0] + (a.y * g.x - a.x * g.y) * n[1] + (a.x * b.y - a.y * b.x) * n[2]) * inv;
        normal[3 * i] = (s0 + 1.f) * 0.5f, normal[3 * i + 1] = (s1 + 1.f) * 0.5f, normal[3 * i + 2] = (s2 + 1.f) * 0.5f;
    }
}

__global__ void __launch_bounds__(256) k_generate_rays(int64_t n, int cam, int W, const int64_t* __restrict__ indices, const float* __restrict__ intr,
                                                       const float* __restrict__ ext, int64_t* __restrict__ cam_idx, float* __restrict__ origins,
                                                       float* __restrict__ directions, float* __restrict__ dnorm, float* __restrict__ area) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t c, py, px;
    if (indices) {
        c = indices[3 * i], py = indices[3 * i + 1], px = indices[3 * i + 2];
    } else {
        c = cam, py = i / W, px = i - py * W;
    }
    float o[3], d[3], nrm, ar;
    pinhole_ray(intr, ext, c, py, px, o, d, &nrm, &ar);
    store3(origins, i, o);
    store3(directions, i, d);
    dnorm[i] = nrm;
    area[i] = ar;
    cam_idx[i] = c;
}

__global__ void __launch_bounds__(128) k_pose_exp_map(int64_t n, int mode, const float* __restrict__ tangent, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float t[6], M[3][4];
#pragma unroll
    for (int k = 0; k < 6; ++k) t[k] = tangent[6 * i + k];
    exp_map<float>(mode, t, M);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) out[12 * i + 4 * a + b] = M[a][b];
}

// per-camera cotangent of [R|t]: G[c][a][b] += d_dir[a] * dir_raw[b] (b<3), G[c][a][3] += d_origin[a].
// Rays of one batch hit K <= a few hundred cameras: warp-level match on the camera index merges equal keys before the reductions.
__global__ void __launch_bounds__(128) k_pose_cotangent(int64_t B, const int64_t* __restrict__ cam_idx, const float* __restrict__ draw,
                                                        const float* __restrict__ d_o, const float* __restrict__ d_d, float* __restrict__ G) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < B;
    const int c = live ? (int)cam_idx[i] : -1;
    float v[12];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float g = live ? d_d[3 * i + a] : 0.f;
#pragma unroll
        for (int b = 0; b < 3; ++b) v[4 * a + b] = live ? g * draw[3 * i + b] : 0.f;
        v[4 * a + 3] = live ? d_o[3 * i + a] : 0.f;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    // reduce within the peer group: every member sums over the group's lanes by broadcast
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        float s = 0.f;
        for (unsigned m = peers; m; m &= m - 1) s += __shfl_sync(peers, v[k], __ffs(m) - 1);
        v[k] = s;
    }
    if (live && lane == leader) {
#pragma unroll
        for (int k = 0; k < 12; ++k) atomicAdd(G + 12 * c + k, v[k]);
    }
}

// d_pose[c][k] += < G[c], d exp_map(pose[c]) / d t_k >  (forward-mode derivative of the same exp_map the forward uses)
__global__ void __launch_bounds__(128) k_pose_grad(int K, int mode, const float* __restrict__ pose, const float* __restrict__ G, float* __restrict__ d_pose) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= K * 6) return;
    const int c = t / 6, k = t - 6 * c;
    Dual tv[6], M[3][4];
#pragma unroll
    for (int j = 0; j < 6; ++j) tv[j] = {pose[6 * c + j], j == k ? 1.f : 0.f};
    exp_map<Dual>(mode, tv, M);
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc += G[12 * c + 4 * a + b] * M[a][b].d;
    d_pose[t] += acc;
}

}  // namespace

extern "C" int nvo_batch_prologue(void* stream, int64_t B, int32_t K, const int32_t* K_dev, int32_t H, int32_t W, const float* u, const float* intrinsics,
                                  const float* extrinsics,
                                  const float* frames_color, const float* frames_depth, const float* frames_normal, const float* pose_adjustment,
                                  int32_t pose_mode, int64_t* indices, int64_t* camera_indices, float* origins, float* directions,
                                  float* directions_norm, float* pixel_area, float* rgb, float* depth, float* normal, float* directions_raw) {
    NVO_CHECK(B >= 0 && K > 0 && H > 0 && W > 0, "nvo_batch_prologue: bad sizes B=%lld K=%d H=%d W=%d", (long long)B, K, H, W);
    NVO_CHECK(pose_mode >= NVO_POSE_OFF && pose_mode <= NVO_POSE_SE3, "nvo_batch_prologue: unknown pose_mode %d", pose_mode);
    NVO_CHECK(pose_mode == NVO_POSE_OFF || pose_adjustment, "nvo_batch_prologue: pose_mode %d needs pose_adjustment", pose_mode);
    if (B == 0) return 0;
    NVO_CHECK(u && intrinsics && extrinsics && frames_color && frames_depth, "nvo_batch_prologue: null input");
    NVO_CHECK(indices && camera_indices && origins && directions && directions_norm && pixel_area && rgb && depth, "nvo_batch_prologue: null output");
    NVO_CHECK(!frames_normal || normal, "nvo_batch_prologue: frames_normal given without a normal output");
    NVO_CHECK((((uintptr_t)intrinsics | (uintptr_t)extrinsics) & 15) == 0, "nvo_batch_prologue: intrinsics / extrinsics must be 16-byte aligned");
    if (B == 0) return 0;
    k_batch_prologue<<<nvo_blocks(B, 128), 128, 0, (cudaStream_t)stream>>>(B, K, K_dev, H, W, u, intrinsics, extrinsics, frames_color, frames_depth, frames_normal,
                                                                          pose_adjustment, pose_mode, indices, camera_indices, origins, directions,
                                                                          directions_norm, pixel_area, rgb, depth, normal, directions_raw);
    NVO_CUDA_LAUNCH_CHECK("k_batch_prologue");
    return 0;
}

extern "C" int nvo_generate_rays(void* stream, int64_t n, int32_t cam, int32_t W, const int64_t* indices, const float* intrinsics, const float* extrinsics,
                                 int64_t* camera_indices, float* origins, float* directions, float* directions_norm, float* pixel_area) {
    NVO_CHECK(n >= 0 && (indices || (W > 0 && cam >= 0)), "nvo_generate_rays: bad arguments n=%lld cam=%d W=%d", (long long)n, cam, W);
    NVO_CHECK(intrinsics && extrinsics && camera_indices && origins && directions && directions_norm && pixel_area, "nvo_generate_rays: null pointer");
    NVO_CHECK((((uintptr_t)intrinsics | (uintptr_t)extrinsics) & 15) == 0, "nvo_generate_rays: intrinsics / extrinsics must be 16-byte aligned");
    if (n == 0) return 0;
    k_generate_rays<<<nvo_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(n, cam, W, indices, intrinsics, extrinsics, camera_indices, origins, directions,
                                                                         directions_norm, pixel_area);
    NVO_CUDA_LAUNCH_CHECK("k_generate_rays");
    return 0;
}

extern "C" int nvo_pose_exp_map(void* stream, int64_t n, int32_t pose_mode, const float* tangent, float* matrices) {
    NVO_CHECK(pose_mode == NVO_POSE_SO3XR3 || pose_mode == NVO_POSE_SE3, "nvo_pose_exp_map: pose_mode must be SO3xR3 (1) or SE3 (2), got %d", pose_mode);
    NVO_CHECK(n >= 0 && tangent && matrices, "nvo_pose_exp_map: bad arguments");
    if (n == 0) return 0;
    k_pose_exp_map<<<nvo_blocks(n, 128), 128, 0, (cudaStream_t)stream>>>(n, pose_mode, tangent, matrices);
    NVO_CUDA_LAUNCH_CHECK("k_pose_exp_map");
    return 0;
}

// CameraOptimizer.apply_to_raybundle on an existing bundle (NS/cameras/camera_optimizers.py:142-147) in one launch: every ray evaluates its
// camera's exp map (a few dozen flops) instead of the gather / add / bmm chain over a [K,3,4] matrix table
__global__ void __launch_bounds__(128) k_pose_apply(int64_t B, int mode, const int64_t* __restrict__ cam_idx, const float* __restrict__ pose,
                                                    const float* __restrict__ o, const float* __restrict__ d, float* __restrict__ oo, float* __restrict__ od) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float* tp = pose + 6 * __ldg(cam_idx + i);
    float t[6], M[3][4];
#pragma unroll
    for (int k = 0; k < 6; ++k) t[k] = __ldg(tp + k);
    exp_map<float>(mode, t, M);
    const float dv[3] = {__ldg(d + 3 * i), __ldg(d + 3 * i + 1), __ldg(d + 3 * i + 2)};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        oo[3 * i + a] = __ldg(o + 3 * i + a) + M[a][3];
        od[3 * i + a] = M[a][0] * dv[0] + M[a][1] * dv[1] + M[a][2] * dv[2];
    }
}

extern "C" int nvo_pose_apply(void* stream, int64_t B, int32_t pose_mode, const int64_t* camera_indices, const float* pose_adjustment, const float* origins,
                              const float* directions, float* origins_out, float* directions_out) {
    NVO_CHECK(pose_mode == NVO_POSE_SO3XR3 || pose_mode == NVO_POSE_SE3, "nvo_pose_apply: pose_mode must be SO3xR3 (1) or SE3 (2), got %d", pose_mode);
    NVO_CHECK(B >= 0, "nvo_pose_apply: negative size");
    if (B == 0) return 0;
    NVO_CHECK(camera_indices && pose_adjustment && origins && directions && origins_out && directions_out, "nvo_pose_apply: null pointer");
    k_pose_apply<<<nvo_blocks(B, 128), 128, 0, (cudaStream_t)stream>>>(B, pose_mode, camera_indices, pose_adjustment, origins, directions, origins_out,
                                                                        directions_out);
    NVO_CUDA_LAUNCH_CHECK("k_pose_apply");
    return 0;
}

extern "C" int nvo_pose_correction_backward(void* stream, int64_t B, int32_t K, int32_t pose_mode, const int64_t* camera_indices, const float* directions_raw,
                                            const float* d_origins, const float* d_directions, const float* pose_adjustment, float* scratch,
                                            float* d_pose) {
    NVO_CHECK(pose_mode == NVO_POSE_SO3XR3 || pose_mode == NVO_POSE_SE3, "nvo_pose_correction_backward: pose_mode must be SO3xR3 (1) or SE3 (2), got %d", pose_mode);
    NVO_CHECK(B >= 0 && K > 0 && pose_adjustment && scratch && d_pose, "nvo_pose_correction_backward: bad arguments");
    NVO_CHECK(B == 0 || (camera_indices && directions_raw && d_origins && d_directions), "nvo_pose_correction_backward: null ray buffers with B > 0");
    if (B > 0) {
        k_pose_cotangent<<<nvo_blocks(B, 128), 128, 0, (cudaStream_t)stream>>>(B, camera_indices, directions_raw, d_origins, d_directions, scratch);
        NVO_CUDA_LAUNCH_CHECK("k_pose_cotangent");
    }
    k_pose_grad<<<nvo_blocks((int64_t)K * 6, 128), 128, 0, (cudaStream_t)stream>>>(K, pose_mode, pose_adjustment, scratch, d_pose);
    NVO_CUDA_LAUNCH_CHECK("k_pose_grad");
    return 0;
}
