// Evaluation frame output: on-device conversion of a rendered frame to the reference's file formats and the
// masked reduction of the depth-scale alignment pass.  reference: evaluation/nerf_renderer.py:160-167, evaluation/renderer.py:79-97,113-121.
#include "nvo_common.cuh"

namespace {

// 4 pixels per thread: 12 colour bytes leave as three 32-bit stores, depth as one float4
__global__ void __launch_bounds__(256) k_frame_finalize(int64_t n, const float* __restrict__ rgb, const float* __restrict__ depth,
                                                        const float* __restrict__ dnorm, float sa, float sb, uint8_t* __restrict__ color,
                                                        float* __restrict__ dout, uint16_t* __restrict__ d16) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i0 = 4 * q;
    if (i0 >= n) return;
    if (i0 + 4 <= n) {
        const float4* r4 = reinterpret_cast<const float4*>(rgb + 3 * i0);
        const float4 a = __ldg(r4), b = __ldg(r4 + 1), c = __ldg(r4 + 2);
        const float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        uint32_t w[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            uint32_t word = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) word |= (uint32_t)(uint8_t)(int)__fmul_rn(v[4 * k + j], 255.f) << (8 * j);
            w[k] = word;
        }
        uint32_t* cw = reinterpret_cast<uint32_t*>(color + 3 * i0);
        cw[0] = w[0], cw[1] = w[1], cw[2] = w[2];
        float4 d = __ldg(reinterpret_cast<const float4*>(depth + i0));
        if (dnorm) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(dnorm + i0));
            d = make_float4(__fdiv_rn(d.x, m.x), __fdiv_rn(d.y, m.y), __fdiv_rn(d.z, m.z), __fdiv_rn(d.w, m.w));
        }
        *reinterpret_cast<float4*>(dout + i0) = d;
        if (d16) {
            const float e[4] = {d.x, d.y, d.z, d.w};
            uint16_t h[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = (uint16_t)(int)__fmul_rn(__fmul_rn(e[j], sa), sb);
            *reinterpret_cast<uint2*>(d16 + i0) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
        }
        return;
    }
    for (int64_t i = i0; i < n; ++i) {  // ragged tail
#pragma unroll
        for (int k = 0; k < 3; ++k) color[3 * i + k] = (uint8_t)(int)__fmul_rn(rgb[3 * i + k], 255.f);
        const float d = dnorm ? __fdiv_rn(depth[i], dnorm[i]) : depth[i];
        dout[i] = d;
        if (d16) d16[i] = (uint16_t)(int)__fmul_rn(__fmul_rn(d, sa), sb);
    }
}

__global__ void __launch_bounds__(256) k_depth_scale_sums(int64_t n, const float* __restrict__ gt, const float* __restrict__ pred, double* __restrict__ sums) {
    double sg = 0.0, sp = 0.0, cnt = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = __ldg(gt + i), p = __ldg(pred + i);
        if (g > 0.f && p > 0.f && g < 5.f && p < 5.f) sg += (double)g, sp += (double)p, cnt += 1.0;
    }
    sg = nvo_warp_sum(sg), sp = nvo_warp_sum(sp), cnt = nvo_warp_sum(cnt);
    __shared__ double sh[3][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[0][wid] = sg, sh[1][wid] = sp, sh[2][wid] = cnt;
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[threadIdx.x][k];
        atomicAdd(sums + threadIdx.x, t);
    }
}

// Point-cloud export (evaluation/nerf_renderer.py:170-209 -> NS/exporter/exporter_utils.py:130-180): per ray the surface point
// origin + direction * depth, the opacity (> 0.5) and axis-aligned box tests, the decoded normal n * 2 - 1 and its re-orientation against
// the view direction (exporter_utils.py:222-226: flipped where dot(view, n) > 0).  keep[i] = 1 for the rays the reference's two masks keep.
__global__ void __launch_bounds__(256) k_point_cloud(int64_t n, const float* __restrict__ origins, const float* __restrict__ directions,
                                                     const float* __restrict__ depth, const float* __restrict__ accumulation,
                                                     const float* __restrict__ normals_coded, float3 lo, float3 hi, int use_box, int reorient,
                                                     float* __restrict__ points, float* __restrict__ normals, uint8_t* __restrict__ keep) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d = __ldg(depth + i);
    float pt[3], v[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        v[a] = __ldg(directions + 3 * i + a);
        pt[a] = __fadd_rn(__ldg(origins + 3 * i + a), __fmul_rn(v[a], d));  // torch: origins + directions * depth
        points[3 * i + a] = pt[a];
    }
    bool k = __ldg(accumulation + i) > 0.5f;
    if (use_box) k = k && pt[0] > lo.x && pt[1] > lo.y && pt[2] > lo.z && pt[0] < hi.x && pt[1] < hi.y && pt[2] < hi.z;
    keep[i] = k ? 1 : 0;
    if (normals_coded) {
        float nn[3], dot = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            nn[a] = __fsub_rn(__fmul_rn(__ldg(normals_coded + 3 * i + a), 2.f), 1.f);
            dot = __fadd_rn(dot, __fmul_rn(v[a], nn[a]));
        }
        const float sgn = (reorient && dot > 0.f) ? -1.f : 1.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) normals[3 * i + a] = nn[a] * sgn;
    }
}

}  // namespace

extern "C" int nvo_frame_finalize(void* stream, int64_t n, const float* rgb, const float* depth, const float* directions_norm, float scale_a, float scale_b,
                                  void* color, float* depth_out, void* depth16) {
    NVO_CHECK(n >= 0, "nvo_frame_finalize: negative n");
    if (n == 0) return 0;
    NVO_CHECK(rgb && depth && color && depth_out, "nvo_frame_finalize: null pointer");
    NVO_CHECK((((uintptr_t)rgb | (uintptr_t)depth | (uintptr_t)directions_norm | (uintptr_t)depth_out) & 15) == 0 && ((uintptr_t)color & 3) == 0 &&
                  ((uintptr_t)depth16 & 7) == 0,
              "nvo_frame_finalize: buffers must be 16-byte (float), 4-byte (colour), 8-byte (uint16 depth) aligned");
    k_frame_finalize<<<nvo_blocks((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(n, rgb, depth, directions_norm, scale_a, scale_b, (uint8_t*)color, depth_out,
                                                                                    (uint16_t*)depth16);
    NVO_CUDA_LAUNCH_CHECK("k_frame_finalize");
    return 0;
}

extern "C" int nvo_depth_scale_sums(void* stream, int64_t n, const float* depth_gt, const float* depth_pred, void* sums) {
    NVO_CHECK(n >= 0 && sums, "nvo_depth_scale_sums: bad arguments");
    if (n == 0) return 0;
    NVO_CHECK(depth_gt && depth_pred, "nvo_depth_scale_sums: null pointer");
    const int64_t want = (n + 255) / 256;
    const unsigned blocks = (unsigned)(want < 4 * nvo_sm_count() ? want : 4 * nvo_sm_count());
    k_depth_scale_sums<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, depth_gt, depth_pred, (double*)sums);
    NVO_CUDA_LAUNCH_CHECK("k_depth_scale_sums");
    return 0;
}

extern "C" int nvo_point_cloud(void* stream, int64_t n, const float* origins, const float* directions, const float* depth, const float* accumulation,
                               const float* normals_coded, const float* box_min, const float* box_max, int32_t reorient, float* points, float* normals,
                               void* keep) {
    NVO_CHECK(n >= 0, "nvo_point_cloud: negative n");
    if (n == 0) return 0;
    NVO_CHECK(origins && directions && depth && accumulation && points && keep, "nvo_point_cloud: null pointer");
    NVO_CHECK(!normals_coded || normals, "nvo_point_cloud: normals output missing");
    NVO_CHECK((box_min == nullptr) == (box_max == nullptr), "nvo_point_cloud: box_min and box_max go together");
    float3 lo = make_float3(0.f, 0.f, 0.f), hi = lo;
    if (box_min) {  // HOST pointers: three floats each
        lo = make_float3(box_min[0], box_min[1], box_min[2]);
        hi = make_float3(box_max[0], box_max[1], box_max[2]);
        NVO_CHECK(lo.x < hi.x && lo.y < hi.y && lo.z < hi.z, "nvo_point_cloud: bounding box min must be smaller than max");
    }
    k_point_cloud<<<nvo_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(n, origins, directions, depth, accumulation, normals_coded, lo, hi, box_min != nullptr,
                                                                       reorient, points, normals, (uint8_t*)keep);
    NVO_CUDA_LAUNCH_CHECK("k_point_cloud");
    return 0;
}
