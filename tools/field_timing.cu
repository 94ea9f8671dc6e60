// Standalone probe (not product code): phase timestamps of the fused field forward kernel (csrc/field_tc.cu built with -DNVO_FT_TIMING).
// Random finite inputs of the config-2 shape; prints, per tile of CTA 0 / group 0, the cycles spent in each wait / epilogue / barrier
// for an epilogue warp of each column half (tid 0: hf 0, tid 128: hf 1).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I nerf-vo_b200/csrc -o /tmp/field_timing tools/field_timing.cu
#define NVO_FT_TIMING 1
#include "../nerf-vo_b200/csrc/api.cu"
#include "../nerf-vo_b200/csrc/field_tc.cu"
#include <vector>

__global__ void fill_half(__half* p, size_t n, float scale, unsigned seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u + seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        p[i] = __float2half(((h & 0xffff) / 65536.f - 0.5f) * scale);
    }
}
__global__ void fill_float(float* p, size_t n, float scale, float offset, unsigned seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u + seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        p[i] = ((h & 0xffff) / 65536.f - 0.5f) * scale + offset;
    }
}

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    const int64_t B = argc > 1 ? atoll(argv[1]) : 4096;
    const int S = 48;
    const int save = argc > 2 ? atoi(argv[2]) : 1;
    const int64_t n = B * S, tiles = (n + 127) / 128;
    __half *feat, *jac;
    float *pos, *dirs, *emb, *sel, *density, *rgb, *pn, *normals, *h0, *pn_raw, *base, *head, *pnp;
    int64_t* cam;
    unsigned char *wimg, *saved;
    cudaMalloc(&feat, tiles * 128 * 32 * 2);
    cudaMalloc(&jac, tiles * 128 * 96 * 2);
    cudaMalloc(&pos, n * 12); cudaMalloc(&dirs, B * 12); cudaMalloc(&emb, 64 * 32 * 4); cudaMalloc(&sel, n * 4);
    cudaMalloc(&density, n * 4); cudaMalloc(&rgb, n * 12); cudaMalloc(&pn, n * 12); cudaMalloc(&normals, n * 12); cudaMalloc(&h0, n * 4);
    cudaMalloc(&pn_raw, n * 12); cudaMalloc(&cam, B * 8);
    cudaMalloc(&base, 4096 * 4); cudaMalloc(&head, 16384 * 4); cudaMalloc(&pnp, 16384 * 4);
    cudaMalloc(&wimg, nvo_field_wimage_bytes()); cudaMalloc(&saved, nvo_field_saved_bytes(n, 0));
    fill_half<<<1024, 256>>>(feat, tiles * 128 * 32, 0.4f, 1);
    fill_half<<<1024, 256>>>(jac, tiles * 128 * 96, 0.4f, 2);
    fill_float<<<1024, 256>>>(pos, n * 3, 2.f, 0.f, 3);
    fill_float<<<64, 256>>>(dirs, B * 3, 1.f, 0.f, 4);
    fill_float<<<8, 256>>>(emb, 64 * 32, 0.2f, 0.f, 5);
    fill_float<<<1024, 256>>>(sel, n, 0.f, 1.f, 6);
    fill_float<<<16, 256>>>(base, 4096, 0.3f, 0.f, 7);
    fill_float<<<64, 256>>>(head, 16384, 0.3f, 0.f, 8);
    fill_float<<<64, 256>>>(pnp, 16384, 0.3f, 0.f, 9);
    cudaMemset(cam, 0, B * 8);
    nvo_grid_desc gd = {};
    gd.n_levels = 16, gd.log2_T = 19;
    for (int i = 0; i < 16; ++i) gd.scalings[i] = 16.f * powf(1.38f, (float)i);
    if (nvo_field_pack_weights(0, &gd, base, head, pnp, wimg)) { printf("pack: %s\n", nvo_last_error()); return 1; }
    unsigned long long* tim;
    cudaMalloc(&tim, 2 * FT_MAX_MARKS * 8);
    cudaMemset(tim, 0, 2 * FT_MAX_MARKS * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        unsigned long long* arg = rep == 2 ? tim : nullptr;
        cudaMemcpyToSymbol(g_ft_timing, &arg, sizeof(arg));
        cudaEventRecord(e0);
        if (nvo_field_forward(0, B, S, feat, jac, pos, dirs, cam, emb, sel, wimg, density, rgb, pn, normals, h0, pn_raw, save ? saved : nullptr, 0)) {
            printf("forward: %s\n", nvo_last_error());
            return 1;
        }
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(e)); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("rep %d: %.1f us (B=%lld save=%d)\n", rep, ms * 1e3, (long long)B, save);
    }
    // ---- backward: phase timestamps of group 0 and of the issuer's view of group 0 (needs the saved tiles of a forward with save=1) ----
    if (save && argc > 3 && atoi(argv[3])) {
        float *ddensity, *drgb, *scratch, *dfeat, *dbase, *dhead, *demb;
        cudaMalloc(&ddensity, n * 4); cudaMalloc(&drgb, n * 12); cudaMalloc(&scratch, 8); cudaMalloc(&dfeat, tiles * 128 * 32 * 4);
        cudaMalloc(&dbase, 4096 * 4); cudaMalloc(&dhead, 16384 * 4); cudaMalloc(&demb, 64 * 32 * 4);
        fill_float<<<1024, 256>>>(ddensity, n, 1e-3f, 0.f, 11);
        fill_float<<<1024, 256>>>(drgb, n * 3, 1e-3f, 0.f, 12);
        cudaMemset(dbase, 0, 4096 * 4); cudaMemset(dhead, 0, 16384 * 4); cudaMemset(demb, 0, 64 * 32 * 4);
        unsigned long long* tb;
        cudaMalloc(&tb, 2 * FT_MAX_MARKS * 8);
        cudaMemset(tb, 0, 2 * FT_MAX_MARKS * 8);
        const int skip = argc > 4 ? atoi(argv[4]) : 0;
        cudaMemcpyToSymbol(g_fb_skip, &skip, sizeof(skip));
        if (skip) printf("TIMING EXPERIMENT: skip bits %d (results invalid)\n", skip);
        for (int rep = 0; rep < 3; ++rep) {
            unsigned long long* arg = rep == 2 ? tb : nullptr;
            cudaMemcpyToSymbol(g_fb_timing, &arg, sizeof(arg));
            cudaEventRecord(e0);
            if (nvo_field_backward(0, B, S, feat, saved, 0, wimg, rgb, h0, sel, cam, ddensity, drgb, nullptr, scratch, dfeat, dbase, dhead, demb, nullptr, nullptr)) {
                printf("backward: %s\n", nvo_last_error());
                return 1;
            }
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(e)); return 1; }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("backward rep %d: %.1f us (absmax pass included)\n", rep, ms * 1e3);
        }
        {
            // per-CTA wall-clock phases (%globaltimer) of one more launch
            unsigned long long* ct;
            cudaMalloc(&ct, 148 * 8 * 8);
            cudaMemset(ct, 0, 148 * 8 * 8);
            cudaMemcpyToSymbol(g_fb_cta, &ct, sizeof(ct));
            unsigned long long* none = nullptr;
            cudaMemcpyToSymbol(g_fb_timing, &none, sizeof(none));
            nvo_field_backward(0, B, S, feat, saved, 0, wimg, rgb, h0, sel, cam, ddensity, drgb, nullptr, scratch, dfeat, dbase, dhead, demb, nullptr, nullptr);
            cudaDeviceSynchronize();
            std::vector<unsigned long long> hc(148 * 8);
            cudaMemcpy(hc.data(), ct, hc.size() * 8, cudaMemcpyDeviceToHost);
            unsigned long long t0 = ~0ULL;
            for (int c = 0; c < 148; ++c) if (hc[8 * c] && hc[8 * c] < t0) t0 = hc[8 * c];
            printf("---- per CTA, ns since the first CTA's entry: entry | setup done | tile loop done | flushed\n");
            double s1 = 0, s2 = 0, s3 = 0, mx2 = 0, mx3 = 0, mn2 = 1e18;
            int nc = 0;
            for (int c = 0; c < 148; ++c) {
                if (!hc[8 * c]) continue;
                ++nc;
                const double a = hc[8 * c] - t0, b = hc[8 * c + 1] - t0, d = hc[8 * c + 2] - t0, e = hc[8 * c + 3] - t0;
                if (c < 6 || c % 37 == 0) printf("  CTA %3d: %7.0f %7.0f %7.0f %7.0f\n", c, a, b, d, e);
                s1 += b, s2 += d, s3 += e;
                if (d > mx2) mx2 = d;
                if (d < mn2) mn2 = d;
                if (e > mx3) mx3 = e;
            }
            printf("  mean setup done %.0f, tile loop done mean %.0f min %.0f max %.0f, flushed mean %.0f max %.0f ns (%d CTAs)\n", s1 / nc, s2 / nc, mn2, mx2, s3 / nc, mx3, nc);
        }
        std::vector<unsigned long long> hb(2 * FT_MAX_MARKS);
        cudaMemcpy(hb.data(), tb, hb.size() * 8, cudaMemcpyDeviceToHost);
        const unsigned long long *gm = hb.data(), *im = hb.data() + FT_MAX_MARKS;
        // group marks per step: epilogue-done, arrived, MMAs-retired; issuer marks per step: block reached, ready observed, load observed, issued
        const char* bs[5] = {"head L2", "head L1", "head L0", "base L1", "base L0"};
        printf("---- backward, CTA 0 group 0: cycles per step\n");
        for (int q = 0; q < 40 && gm[3 * q + 2] && im[4 * q + 3]; ++q) {
            const unsigned long long epi_done = gm[3 * q], arrived = gm[3 * q + 1], retired = gm[3 * q + 2], at_block = im[4 * q], ready_seen = im[4 * q + 1], seen = im[4 * q + 2], issued = im[4 * q + 3];
            const unsigned long long next = gm[3 * (q + 1)];
            printf("  tile %d %-8s fence+arrive %5llu | issuer reaches block %6lld | waits ready %5llu | waits load %5llu | issue %5llu | MMA + commit %5llu | epilogue %5llu | step total %6llu\n", q / 5,
                   bs[q % 5], arrived - epi_done, (long long)(at_block - arrived), ready_seen - at_block, seen - ready_seen, issued - seen, retired - issued, next ? next - retired : 0ULL,
                   next ? next - epi_done : 0ULL);
        }
        return 0;
    }
    std::vector<unsigned long long> h(2 * FT_MAX_MARKS);
    cudaMemcpy(h.data(), tim, h.size() * 8, cudaMemcpyDeviceToHost);
    // marks per tile: [in-wait pre] then per step: commit-wait pre, post, sync pre, sync post
    const char* steps[9] = {"S0 base0", "S1 base1+ns", "S2 head0", "S3 head1", "S4 head2+pn0", "S5 pn1", "S6 pn2", "S7 pn3", "-"};
    for (int th = 0; th < 2; ++th) {
        const unsigned long long* t = h.data() + th * FT_MAX_MARKS;
        printf("---- thread %d (hf %d)\n", th * 128, th);
        int m = 0;
        for (int tile = 0; tile < 6 && t[m]; ++tile) {
            const unsigned long long tile0 = t[m];
            printf("tile %d: start +%llu since kernel mark0\n", tile, tile0 - t[0]);
            m += 1;
            for (int s = 0; s < 8; ++s) {
                if (!t[m + 3]) break;
                printf("  %-12s issue->wait %5llu | mma wait %5llu | epilogue %5llu | sync %5llu\n", steps[s], t[m] - (s == 0 ? tile0 : t[m - 1]), t[m + 1] - t[m],
                       t[m + 2] - t[m + 1], t[m + 3] - t[m + 2]);
                m += 4;
            }
            printf("  tile total %llu cycles\n", t[m - 1] - tile0);
        }
    }
    return 0;
}
