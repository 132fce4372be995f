/* epipolar_b200.h — C ABI of the B200-native epipolar-transformer fusion path.
 *
 * Drop-in boundary.  The reference has no FFI: its "operator" is the
 * Python call  Epipolar.forward(feat1, feat2, P1, P2, ...)  at
 *   /root/reference/modeling/layers/epipolar.py:82          (the forward being replaced)
 *   /root/reference/modeling/layers/epipolar.py:323-418     (grid2sample_locs — fused in-kernel)
 *   /root/reference/modeling/layers/epipolar.py:272-321     (epipolar_similarity — fused)
 *   /root/reference/modeling/layers/epipolar.py:248-255     (z conv + BN + residual — epilogue)
 *   /root/reference/vision/multiview.py:16-21,25-57,154-163 (camera_center / normalize /
 *                                                            de_normalize / pix2coord / coord2pix)
 *   /root/reference/modeling/backbones/resnet.py:385-388    (caller residual `ret + feat`)
 * A maintainer binds these symbols with ctypes; the host side shipped
 * here (epipolar_transformers_b200/epipolar.py) does exactly that.
 *
 * Conventions: plain pointers + sizes, no torch types.  All tensor pointers are DEVICE
 * pointers to float32 unless a name ends in _host.  The library never allocates persistent
 * device memory; the caller passes a workspace.  Every entry point is re-entrant, takes the
 * CUDA stream explicitly, never synchronises the device, and returns 0 on success or a
 * negative EPI_E* code (epi_last_error() gives a thread-local message).
 */
#ifndef EPIPOLAR_B200_H_
#define EPIPOLAR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EPI_ABI_VERSION 2

#define EPI_OK 0
#define EPI_EINVAL (-1)       /* bad argument / unsupported shape */
#define EPI_EWORKSPACE (-2)   /* workspace too small */
#define EPI_ECUDA (-3)        /* CUDA runtime error at launch (message has the cudaError string) */

/* kernel variants (EpiFusionParams.variant) */
#define EPI_VARIANT_AUTO 0    /* pipelined kernel when the shape allows, else sector / block tiles, else warp */
#define EPI_VARIANT_WARP 1    /* one warp per reference pixel, online softmax (baseline kernel) */
#define EPI_VARIANT_TILE 2    /* tensor-core kernel, 4x8 pixel tiles: shared-memory staged source taps, score interpolation */
#define EPI_VARIANT_SECTOR 3  /* tensor-core kernel, tiles of 32 pixels that share an epipolar line (sorted by epipolar angle) */
#define EPI_VARIANT_PIPE 4    /* warp-specialised, mbarrier-pipelined tensor-core kernel over epipolar-sector work items (default) */

typedef struct EpiFusionParams {
    /* ---- inputs ---------------------------------------------------------------------- */
    const float *feat_ref;        /* [N,C,H,W] logical; element strides below (NCHW or channels-last) */
    int64_t ref_stride[4];
    const float *feat_src;        /* [N,C,H,W] logical */
    int64_t src_stride[4];
    const float *P_ref;           /* [N,3,4] contiguous: KRT of the reference view  (forward arg P1) */
    const float *P_src;           /* [N,3,4] contiguous: KRT of the source view     (forward arg P2) */
    const float *sample_locs_in;  /* optional [K,N,H,W,2] normalised grid coords: replaces the fused
                                     geometry (parity protocol T1: inject the reference's own locations) */
    /* ---- outputs --------------------------------------------------------------------- */
    float *out;                   /* [N,C,H,W] logical, strides below */
    int64_t out_stride[4];
    float *attn;                  /* optional [N,K,H,W] contiguous: softmax weights ("depth", epipolar.py:263) */
    float *corr_pos;              /* optional [N,H,W,2] contiguous: arg-max correspondence, feature px (:237-242) */
    float *sample_locs_out;       /* optional [K,N,H,W,2] contiguous: locations actually sampled (:183) */
    /* ---- optional folded eval-mode epilogue:  y = Wf·o + bf  (+ o if z_residual) ------ */
    const float *z_weight_folded; /* [C,C] row-major (out_ch, in_ch) = diag(gamma/sqrt(var+eps))·Wz, or NULL */
    const float *z_bias_folded;   /* [C] */
    /* ---- scratch ---------------------------------------------------------------------- */
    void *workspace;
    size_t workspace_bytes;       /* >= epi_fusion_workspace_bytes(p) */
    /* ---- shape & semantics ------------------------------------------------------------ */
    int32_t N, C, H, W, K;
    float downsample;             /* cfg.BACKBONE.DOWNSAMPLE */
    float img_scale;              /* cfg.DATASETS.IMAGE_RESIZE * PREDICT_RESIZE */
    float eps;                    /* 1e-3, epipolar.py:20 */
    float softmax_scale;          /* cfg.EPIPOLAR.SOFTMAXSCALE (0.125) */
    int32_t align_corners;        /* grid_sample semantics (torch>=1.3 default 0) */
    int32_t correct_normalize;    /* cfg.EPIPOLAR.USE_CORRECT_NORMALIZE */
    int32_t z_residual;           /* cfg.EPIPOLAR.ZRESIDUAL (only with z_weight_folded) */
    int32_t add_ref_residual;     /* 1: also add feat_ref (the caller's `ret + feat`, resnet.py:388) */
    int32_t variant;              /* EPI_VARIANT_* */
    int32_t reserved[3];
    /* ---- optional persistent state (ABI v2) -------------------------------------------- */
    void *cache;                  /* device memory the caller keeps alive ACROSS calls and zero-fills once, or NULL.  Holds the
                                     per-pair constants and the epipolar pixel order keyed by (P_ref, P_src, H, W, downsample,
                                     img_scale): an unchanged camera pair skips their recomputation.  One cache per
                                     (module, device, stream); never share it between concurrently running calls. */
    size_t cache_bytes;           /* >= epi_fusion_cache_bytes(p) when cache != NULL */
} EpiFusionParams;

/* ABI version of the loaded library (== EPI_ABI_VERSION it was built with). */
int epi_version(void);

/* Thread-local description of the last error returned on this thread. */
const char *epi_last_error(void);

/* Bytes of scratch the forward needs for these shapes/strides/flags (0 is possible). */
size_t epi_fusion_workspace_bytes(const EpiFusionParams *p);

/* Bytes of the optional persistent cache for these shapes (0 when the selected kernel keeps no cross-call state). */
size_t epi_fusion_cache_bytes(const EpiFusionParams *p);

/* The fused forward: geometry + K bilinear taps + softmax(QK)·V (+ z/BN epilogue, + residuals).
 * Replaces Epipolar.forward for ATTENTION='avg', SIMILARITY='dot', SOFTMAX_ENABLED.
 * Asynchronous on `stream` (a cudaStream_t); returns launch-time errors only. */
int epi_fusion_forward_f32(const EpiFusionParams *p, void *stream);

/* ---- backward of the fused attention --------------------------------------------------------
 * Replaces autograd through  F.grid_sample x2 / mul / sum / softmax  of /root/reference/modeling/layers/epipolar.py:188-247
 * (called under autograd by /root/reference/engine/trainer.py:72).  Sample locations are constants (:178 torch.no_grad).
 * grad_keys / grad_vals select the OTHER_GRAD members 'other1' / 'other2' (:141-153).  The z conv + BN of training mode
 * stay in PyTorch, so `grad_out` is the gradient w.r.t. the PRE-z fused feature. */
typedef struct EpiFusionBwdParams {
    const float *feat_ref;        /* [N,C,H,W] logical, strides below */
    int64_t ref_stride[4];
    const float *feat_src;
    int64_t src_stride[4];
    const float *P_ref, *P_src;   /* [N,3,4] */
    const float *sample_locs_in;  /* optional, as in the forward */
    const float *attn;            /* [N,K,H,W] contiguous: the forward's attention output */
    const float *grad_out;        /* [N,C,H,W] logical: dL/d(fused feature) */
    int64_t gout_stride[4];
    const float *grad_attn;       /* optional [N,K,H,W] contiguous: dL/d(attention output) */
    float *grad_ref;              /* optional out [N,C,H,W] logical: dL/dfeat_ref */
    int64_t gref_stride[4];
    float *grad_src;              /* optional out [N,C,H,W] logical: dL/dfeat_src (overwritten, not accumulated) */
    int64_t gsrc_stride[4];
    void *workspace;
    size_t workspace_bytes;       /* >= epi_fusion_backward_workspace_bytes(p) */
    int32_t N, C, H, W, K;
    float downsample, img_scale, eps, softmax_scale;
    int32_t align_corners, correct_normalize;
    int32_t grad_keys, grad_vals; /* 'other1' / 'other2' in cfg.EPIPOLAR.OTHER_GRAD */
    int32_t reserved[4];
} EpiFusionBwdParams;

size_t epi_fusion_backward_workspace_bytes(const EpiFusionBwdParams *p);
int epi_fusion_backward_f32(const EpiFusionBwdParams *p, void *stream);

/* Only the geometry: sample locations [K,N,H,W,2] for (P_ref,P_src)  (grid2sample_locs). */
int epi_sample_locs_f32(const float *P_ref, const float *P_src, float *sample_locs_out, int32_t N,
                        int32_t H, int32_t W, int32_t K, float downsample, float img_scale, float eps,
                        int32_t correct_normalize, void *stream);

/* find_tensor_peak_batch for a whole batch (replaces the per-item Python loop at
 * /root/reference/modeling/backbones/resnet.py:423-428 over modeling/backbones/basic_batch.py:17-63):
 * heatmaps [B,J,H,W] contiguous -> locs [B,J,2] (x, y in image coordinates, pix2coord applied) and scores [B,J].
 * radius = cfg.KEYPOINT.SIGMA; threshold 1e-6 in the reference; int_div = 1 reproduces torch < 1.4's integer `index / W`,
 * 0 the true division of current torch (what the reference computes under the torch installed with this library). */
int epi_find_peaks_f32(const float *heatmaps, float *locs, float *scores, int32_t B, int32_t J, int32_t H, int32_t W,
                       float radius, float downsample, float threshold, int32_t int_div, void *stream);

/* Fold conv1x1 z + eval BatchNorm into (Wf, bf) on the device, no host sync:
 *   Wf[o,c] = s[o]·Wz[o,c],  bf[o] = s[o]·(bz[o] − mean[o]) + beta[o],  s = gamma/sqrt(var+bn_eps). */
int epi_fold_z_bn_f32(const float *z_weight, const float *z_bias, const float *bn_weight,
                      const float *bn_bias, const float *bn_mean, const float *bn_var, float bn_eps,
                      int32_t C, float *w_folded, float *b_folded, void *stream);

/* Diagnostic: one-CTA tcgen05 GEMM in the exact operand forms the fusion kernel uses
 * (mode 0: D[128,N] = A[128,K]·B[N,K]^T, both K-major;  mode 1: D[128,N] = At[K,128]^T·B[N,K]^T, A MN-major;
 * split=1: bf16 (hi,lo) three-term products).  All pointers device fp32, row-major. */
int epi_umma_selftest(int mode, const float *A, const float *B, float *D, int N, int K, int split, void *stream);

/* Measurement aid (bench.py roofline): when enabled on this thread, epi_fusion_forward_f32 brackets its dominant
 * kernel (the fused attention kernel) with CUDA events on the caller's stream; epi_kernel_timing_last_ms()
 * synchronises on them and returns that launch's duration in milliseconds (< 0 if there is none). */
int epi_kernel_timing_enable(int on);
float epi_kernel_timing_last_ms(void);
/* same, per launch group: ms3[0] operand staging, ms3[1] fused attention kernel, ms3[2] epilogue pass */
int epi_kernel_timing_last3(float *ms3);

/* Number of kernels the last successful epi_fusion_forward_f32 on this thread launched. */
int epi_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* EPIPOLAR_B200_H_ */
