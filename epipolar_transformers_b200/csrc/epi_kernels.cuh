// epi_kernels.cuh — internal launch interface between the C ABI (epi_abi.cu) and the kernels.
#pragma once
#include <cuda_bf16.h>

#include "epi_common.cuh"

namespace epi {

// Device-side view of one fused forward (built from EpiFusionParams by the ABI layer).
struct FusionArgs {
    const float *feat_ref;  int64_t ref_stride[4];
    const float *src_nhwc;                    // [N,H,W,C] contiguous, 16-byte aligned (zero-copy or staged) — warp kernel
    const __nv_bfloat16 *src_hi, *src_lo;     // [N,H,W,C] bf16 planes, src ≈ hi + lo — tile kernel
    const __nv_bfloat16 *ref_hi, *ref_lo;     // same for feat_ref — sector tiles only
    const uint16_t *order;                    // [N,H*W] pixels sorted by epipolar angle — sector tiles only (else null)
    const float *P_ref, *P_src;
    const float *locs_in;
    float *out;             int64_t out_stride[4];
    __nv_bfloat16 *out_hi, *out_lo;           // optional: fused feature as bf16 (hi, lo) planes [N,H,W,C] (feeds the z GEMM)
    float *attn, *corr_pos, *locs_out;
    int N, C;
    float softmax_scale;
    int add_ref;
    int *tile_counter;                        // zeroed by the staging kernel; dynamic tile scheduler of the tile kernel
    const PairGeom *pair_geom;                // [N] per-pair constants (fp64-derived) written by the staging kernel — pipe kernel
    int *err_flag;                            // device word OR-ed with 1 when a work item had to be dropped (never for supported shapes)
    uint8_t *plan_cache;                      // optional persistent records of the pipe kernel's work items (camera-only data), or null
    int plan_records;                         // capacity of plan_cache in records of fusion_pipe_plan_record_bytes()
    const uint32_t *pair_epoch;               // [N] stride 32 words: epoch of each pair's cached plan (bumped by the staging kernel on a key miss)
    GeomCfg geom;
};

// Backward of the fused attention (epi_fusion_bwd.cu)
struct BwdArgs {
    const float *feat_ref;  int64_t ref_stride[4];
    const float *src_nhwc;                    // [N,H,W,C] contiguous fp32
    const float *P_ref, *P_src, *locs_in;
    const float *attn;                        // [N,K,H,W] saved by the forward
    const float *grad_out;  int64_t gout_stride[4];
    const float *grad_attn;                   // optional [N,K,H,W]
    float *grad_ref;        int64_t gref_stride[4];   // optional
    float *dsrc_nhwc;                         // optional [N,H,W,C] fp32, zero-initialised accumulator
    int N, C;
    float softmax_scale;
    int grad_keys, grad_vals;
    GeomCfg geom;
};
cudaError_t launch_fusion_bwd(const BwdArgs &a, cudaStream_t st);

// z-projection epilogue:  y[n,o,p] = sum_c Wf[o,c]·x[n,c,p] + bf[o] (+x[n,o,p]) (+ref[n,o,p])
struct ZArgs {
    const float *x;         int64_t x_stride[4];     // pre-z fused feature
    const float *ref;       int64_t ref_stride[4];   // may be null
    float *y;               int64_t y_stride[4];
    const float *Wf, *bf;
    int N, C, HW, W;
    int z_residual, add_ref;
};

// tensor-core z-projection: x arrives as bf16 (hi, lo) planes [N,H,W,C] written by the tile kernel
struct ZGemmArgs {
    const __nv_bfloat16 *x_hi, *x_lo;
    const __nv_bfloat16 *w_hi, *w_lo;         // Wf (+ I when ZRESIDUAL) [C out][C in] as bf16 (hi, lo) planes (written by the staging kernel)
    const float *Wf, *bf;
    const float *ref;       int64_t ref_stride[4];
    float *y;               int64_t y_stride[4];
    int N, C, HW, W, Npad;
    int z_residual, add_ref;
};
bool zgemm_supported(int C);
cudaError_t launch_zgemm(const ZGemmArgs &z, cudaStream_t st);

cudaError_t launch_fusion_warp(const FusionArgs &a, cudaStream_t st);
cudaError_t launch_fusion_tile(const FusionArgs &a, cudaStream_t st);
cudaError_t launch_fusion_pipe(const FusionArgs &a, cudaStream_t st);
bool fusion_pipe_shape_ok(int C, int H, int W, int K, bool has_locs_in);
size_t fusion_pipe_plan_record_bytes();
int fusion_pipe_plan_records(int N, int H, int W);
bool fusion_tile_supported(const FusionArgs &a);
bool fusion_tile_shape_ok(int C, int H, int W, int K, bool has_locs_in);
cudaError_t launch_sector_order(const float *P_ref, const float *P_src, uint16_t *order, int N, const GeomCfg &gc, cudaStream_t st);
cudaError_t launch_split_planes(const float *src, const int64_t stride[4], __nv_bfloat16 *hi, __nv_bfloat16 *lo, int N, int C,
                                int H, int W, int *zero_me, cudaStream_t st);

cudaError_t launch_stage(const float *ref, const int64_t ref_stride[4], const float *src, const int64_t src_stride[4],
                         __nv_bfloat16 *planes, const float *P_ref, const float *P_src, PairGeom *pair_geom, uint16_t *order,
                         float *order_key, const float *Wf, __nv_bfloat16 *w_planes, int w_add_identity, int *zero_words, int N, int C,
                         int H, int W, const GeomCfg &gc, cudaStream_t st);

cudaError_t launch_nchw_to_nhwc(const float *src, const int64_t stride[4], float *dst, int N, int C, int H, int W, cudaStream_t st);
cudaError_t launch_z_epilogue(const ZArgs &z, cudaStream_t st);
cudaError_t launch_unstage(const float *pm, const float *ref, const int64_t ref_stride[4], float *out, const int64_t out_stride[4],
                           int N, int C, int H, int W, cudaStream_t st);
cudaError_t launch_fold_z_bn(const float *zw, const float *zb, const float *g, const float *b, const float *mean,
                             const float *var, float eps, int C, float *wf, float *bf, cudaStream_t st);
cudaError_t launch_peaks(const float *heat, float *locs, float *scores, int B, int J, int H, int W, float radius, float downsample,
                         float threshold, int int_div, cudaStream_t st);
cudaError_t launch_sample_locs(const float *P_ref, const float *P_src, float *locs, int N, const GeomCfg &gc, cudaStream_t st);

}  // namespace epi
