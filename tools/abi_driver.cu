// abi_driver.cu — torch-free driver of the C ABI (include/epipolar_b200.h) for compute-sanitizer:
//   nvcc -o tools/abi_driver tools/abi_driver.cu -Iinclude -Lepipolar_transformers_b200 -lepipolar_b200 -Xlinker -rpath,'$ORIGIN/../epipolar_transformers_b200'
//   compute-sanitizer --tool memcheck tools/abi_driver
// Runs the fused forward (pipe kernel with and without z epilogue, persistent cache: miss then hit), the warp kernel, the backward
// and the peak finder on small synthetic shapes; no PyTorch, no lazy module loading in the way of the sanitizer's report.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "epipolar_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
#define EK(x) do { int r_ = (x); if (r_ != EPI_OK) { printf("ABI error %d (%s) at %s:%d\n", r_, epi_last_error(), __FILE__, __LINE__); exit(3); } } while (0)

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

// H36M-like ring camera v of V (KRT = K [R | -R C]), image size img
static void ring_cam(int v, int V, float img, float *P) {
    const double f = 290.0 * img / 256.0, c = img / 2.0, ang = 2.0 * M_PI * v / V + 0.3;
    const double C[3] = {5000 * cos(ang), 5000 * sin(ang), 1500}, T[3] = {0, 0, 1000};
    double fw[3] = {T[0] - C[0], T[1] - C[1], T[2] - C[2]};
    double n = sqrt(fw[0] * fw[0] + fw[1] * fw[1] + fw[2] * fw[2]);
    for (int i = 0; i < 3; i++) fw[i] /= n;
    double r[3] = {fw[1] * 1 - fw[2] * 0, fw[2] * 0 - fw[0] * 1, 0};     // fwd x up, up = (0,0,1)
    r[0] = fw[1]; r[1] = -fw[0]; r[2] = 0;
    n = sqrt(r[0] * r[0] + r[1] * r[1]);
    for (int i = 0; i < 3; i++) r[i] /= n;
    const double d[3] = {fw[1] * r[2] - fw[2] * r[1], fw[2] * r[0] - fw[0] * r[2], fw[0] * r[1] - fw[1] * r[0]};
    const double R[3][3] = {{r[0], r[1], r[2]}, {d[0], d[1], d[2]}, {fw[0], fw[1], fw[2]}};
    const double Kk[3][3] = {{f, 0, c}, {0, f, c}, {0, 0, 1}};
    double Rt[3][4];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) Rt[i][j] = R[i][j]; Rt[i][3] = -(R[i][0] * C[0] + R[i][1] * C[1] + R[i][2] * C[2]); }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) { double s = 0; for (int k = 0; k < 3; k++) s += Kk[i][k] * Rt[k][j]; P[i * 4 + j] = (float)s; }
}

static float *dev_rand(size_t n, bool relu) {
    std::vector<float> h(n);
    for (size_t i = 0; i < n; i++) { h[i] = frand(); if (relu && h[i] < 0) h[i] = 0; }
    float *d; CK(cudaMalloc(&d, n * 4)); CK(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice));
    return d;
}

static void run_case(int N, int C, int H, int W, int K, int with_z, int variant) {
    const size_t map = (size_t)N * C * H * W;
    float *ref = dev_rand(map, true), *src = dev_rand(map, true), *out, *attn, *corr, *Wf = nullptr, *bf = nullptr;
    CK(cudaMalloc(&out, map * 4)); CK(cudaMalloc(&attn, (size_t)N * K * H * W * 4)); CK(cudaMalloc(&corr, (size_t)N * H * W * 8));
    std::vector<float> P1(N * 12), P2(N * 12);
    for (int n = 0; n < N; n++) { ring_cam(n % 4, 4, 4.f * H, &P1[n * 12]); ring_cam((n + 1) % 4, 4, 4.f * H, &P2[n * 12]); }
    float *dP1, *dP2; CK(cudaMalloc(&dP1, N * 48)); CK(cudaMalloc(&dP2, N * 48));
    CK(cudaMemcpy(dP1, P1.data(), N * 48, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dP2, P2.data(), N * 48, cudaMemcpyHostToDevice));
    if (with_z) { Wf = dev_rand((size_t)C * C, false); bf = dev_rand(C, false); }
    EpiFusionParams p; memset(&p, 0, sizeof(p));
    p.feat_ref = ref; p.feat_src = src; p.P_ref = dP1; p.P_src = dP2; p.out = out; p.attn = attn; p.corr_pos = corr;
    const int64_t st[4] = {(int64_t)C * H * W, (int64_t)H * W, W, 1};
    for (int i = 0; i < 4; i++) { p.ref_stride[i] = st[i]; p.src_stride[i] = st[i]; p.out_stride[i] = st[i]; }
    p.z_weight_folded = Wf; p.z_bias_folded = bf; p.z_residual = with_z;
    p.N = N; p.C = C; p.H = H; p.W = W; p.K = K; p.downsample = 4.f; p.img_scale = 1.f; p.eps = 1e-3f; p.softmax_scale = 0.125f;
    p.correct_normalize = 1; p.variant = variant;
    void *cache = nullptr;
    const size_t cb = epi_fusion_cache_bytes(&p);
    if (cb) { CK(cudaMalloc(&cache, cb)); CK(cudaMemset(cache, 0, cb)); p.cache = cache; p.cache_bytes = cb; }
    const size_t wb = epi_fusion_workspace_bytes(&p);
    void *ws = nullptr; if (wb) CK(cudaMalloc(&ws, wb));
    p.workspace = ws; p.workspace_bytes = wb;
    for (int it = 0; it < 2; it++) { EK(epi_fusion_forward_f32(&p, nullptr)); CK(cudaDeviceSynchronize()); }     // cache miss, then hit
    std::vector<float> h(map); CK(cudaMemcpy(h.data(), out, map * 4, cudaMemcpyDeviceToHost));
    double s = 0; for (size_t i = 0; i < map; i++) s += h[i];
    printf("forward N=%d C=%d %dx%d K=%d z=%d variant=%d launches=%d sum(out)=%.6g\n", N, C, H, W, K, with_z, variant, epi_last_launch_count(), s);
    if (!with_z) {                                                   // backward on the same tensors
        float *gout = dev_rand(map, false), *gref, *gsrc;
        CK(cudaMalloc(&gref, map * 4)); CK(cudaMalloc(&gsrc, map * 4));
        EpiFusionBwdParams b; memset(&b, 0, sizeof(b));
        b.feat_ref = ref; b.feat_src = src; b.P_ref = dP1; b.P_src = dP2; b.attn = attn; b.grad_out = gout; b.grad_ref = gref; b.grad_src = gsrc;
        for (int i = 0; i < 4; i++) { b.ref_stride[i] = st[i]; b.src_stride[i] = st[i]; b.gout_stride[i] = st[i]; b.gref_stride[i] = st[i]; b.gsrc_stride[i] = st[i]; }
        b.N = N; b.C = C; b.H = H; b.W = W; b.K = K; b.downsample = 4.f; b.img_scale = 1.f; b.eps = 1e-3f; b.softmax_scale = 0.125f;
        b.correct_normalize = 1; b.grad_keys = 1; b.grad_vals = 1;
        const size_t bw = epi_fusion_backward_workspace_bytes(&b);
        void *bws; CK(cudaMalloc(&bws, bw)); b.workspace = bws; b.workspace_bytes = bw;
        EK(epi_fusion_backward_f32(&b, nullptr)); CK(cudaDeviceSynchronize());
        printf("  backward ok\n");
        cudaFree(gout); cudaFree(gref); cudaFree(gsrc); cudaFree(bws);
    }
    cudaFree(ref); cudaFree(src); cudaFree(out); cudaFree(attn); cudaFree(corr); cudaFree(dP1); cudaFree(dP2);
    if (Wf) cudaFree(Wf); if (bf) cudaFree(bf); if (ws) cudaFree(ws); if (cache) cudaFree(cache);
}

int main() {
    srand(7);
    run_case(2, 64, 32, 32, 32, 0, EPI_VARIANT_AUTO);
    run_case(2, 64, 32, 32, 32, 1, EPI_VARIANT_AUTO);
    run_case(1, 256, 64, 64, 64, 1, EPI_VARIANT_AUTO);
    run_case(1, 40, 12, 20, 16, 0, EPI_VARIANT_AUTO);       // C % 64 != 0, ragged map
    run_case(1, 512, 24, 24, 32, 1, EPI_VARIANT_PIPE);      // wide: two query-panel halves, single-buffered O, one-tile-per-CTA z GEMM
    run_case(1, 320, 16, 24, 16, 0, EPI_VARIANT_PIPE);      // wide, C % 128 != 0, pixel-major plane + transposition pass
    run_case(1, 16, 136, 136, 16, 0, EPI_VARIANT_PIPE);     // > 16384 pixels: row-windowed union bitmap, pixel order in its own launch
    run_case(1, 24, 16, 16, 16, 0, EPI_VARIANT_WARP);
    run_case(1, 32, 16, 16, 16, 0, EPI_VARIANT_SECTOR);
    // peak finder
    float *heat = dev_rand((size_t)2 * 17 * 64 * 64, true), *locs, *sc;
    CK(cudaMalloc(&locs, 2 * 17 * 2 * 4)); CK(cudaMalloc(&sc, 2 * 17 * 4));
    EK(epi_find_peaks_f32(heat, locs, sc, 2, 17, 64, 64, 8.f, 4.f, 1e-6f, 0, nullptr)); CK(cudaDeviceSynchronize());
    printf("peaks ok\nALL OK\n");
    return 0;
}
