// epi_common.cuh — shared device code of the epipolar fusion kernels (sm_100a).
//
// Geometry restates grid2sample_locs (/root/reference/modeling/layers/epipolar.py:323-418) and
// the helpers of /root/reference/vision/multiview.py (:16-21 camera_center, :25-37 normalize,
// :39-57 de_normalize, :154-163 pix2coord/coord2pix) per pixel, with the better-conditioned
// infinite-homography point x2' = (A2 A1^-1) p on the same epipolar line.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace epi {

constexpr float kFar = 10000.0f;      // epipolar.py:51-53
constexpr float kMasked = -1e10f;     // epipolar.py:298

// Per-(ref,src)-pair constants: M = A2·A1^-1 (row-major 3x3) and the epipole e2/e2.z.
struct PairGeom {
    float M[9];
    float ex, ey;
};

// Launch-invariant geometry configuration.
struct GeomCfg {
    float ds, r, eps;
    float inv_rds, off_ds;          // image coord -> feature px:  pix = v*inv_rds + off_ds
    float gsx, gox, gsy, goy;       // feature px -> grid coord:   g = pix*gs + go   (per axis)
    float xmin, xmax, ymin, ymax;   // image coords of first/last pixel centres (epipolar.py:46-49)
    int correct;                    // USE_CORRECT_NORMALIZE
    int align;                      // grid_sample align_corners
    int H, W, K;
};

__host__ __device__ __forceinline__ float pix2coord(int i, float ds, float r) {
    return ((float)i * ds + ds * 0.5f - 0.5f) * r;           // multiview.py:154-157, epipolar.py:35-38
}

// One thread: fp64 3x3 inverse + products (≈100 flops), so the per-pixel fp32 math starts
// from correctly rounded constants (the reference's fp32 pinv path is noisy).
__device__ inline void pair_geom_from_krt(const float *__restrict__ P1, const float *__restrict__ P2, PairGeom &g) {
    double a[9], b[9], ai[9], t1[3], t2[3];
    for (int r = 0; r < 3; r++) {
        for (int q = 0; q < 3; q++) { a[r * 3 + q] = (double)P1[r * 4 + q]; b[r * 3 + q] = (double)P2[r * 4 + q]; }
        t1[r] = (double)P1[r * 4 + 3]; t2[r] = (double)P2[r * 4 + 3];
    }
    double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    double id = 1.0 / (a[0] * c00 + a[1] * c01 + a[2] * c02);
    ai[0] = c00 * id; ai[1] = (a[2] * a[7] - a[1] * a[8]) * id; ai[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    ai[3] = c01 * id; ai[4] = (a[0] * a[8] - a[2] * a[6]) * id; ai[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    ai[6] = c02 * id; ai[7] = (a[1] * a[6] - a[0] * a[7]) * id; ai[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    double c[3], e[3];
    for (int r = 0; r < 3; r++) c[r] = -(ai[r * 3] * t1[0] + ai[r * 3 + 1] * t1[1] + ai[r * 3 + 2] * t1[2]);   // camera centre
    for (int r = 0; r < 3; r++) {
        e[r] = b[r * 3] * c[0] + b[r * 3 + 1] * c[1] + b[r * 3 + 2] * c[2] + t2[r];                          // epipole P2·[C;1]
        for (int q = 0; q < 3; q++)
            g.M[r * 3 + q] = (float)(b[r * 3] * ai[q] + b[r * 3 + 1] * ai[3 + q] + b[r * 3 + 2] * ai[6 + q]);
    }
    g.ex = (float)(e[0] / e[2]);
    g.ey = (float)(e[1] / e[2]);
}

__device__ __forceinline__ float sdiv(float v, float eps) {       // sign(v)*max(|v|,eps), epipolar.py:370-373
    float a = fmaxf(fabsf(v), eps);
    return v > 0.f ? a : (v < 0.f ? -a : 0.f);
}

// Endpoints (image coords) of the epipolar line of reference pixel (px,py) clipped to the
// pixel-centre rectangle; far sentinel when fewer than two valid intersections (:369-405).
__device__ __forceinline__ void line_endpoints(const PairGeom &g, const GeomCfg &c, float px, float py,
                                               float &sx, float &sy, float &ex, float &ey) {
    float zx = g.M[0] * px + g.M[1] * py + g.M[2];
    float zy = g.M[3] * px + g.M[4] * py + g.M[5];
    float zz = g.M[6] * px + g.M[7] * py + g.M[8];
    float x2 = zx / zz, y2 = zy / zz;
    float l0 = g.ey - y2, l1 = x2 - g.ex, l2 = g.ex * y2 - g.ey * x2;    // e2 × x2, both with z = 1
    float d1 = sdiv(l1, c.eps), d0 = sdiv(l0, c.eps);
    float by1 = -(c.xmin * l0 + l2) / d1;
    float by2 = -(c.xmax * l0 + l2) / d1;
    float bx0 = -(c.ymin * l1 + l2) / d0;
    float bx3 = -(c.ymax * l1 + l2) / d0;
    bool ok0 = (bx0 >= c.xmin + c.eps) && (bx0 < c.xmax - c.eps);
    bool ok1 = (by1 > c.ymin + c.eps) && (by1 <= c.ymax - c.eps);
    bool ok2 = (by2 >= c.ymin + c.eps) && (by2 < c.ymax - c.eps);
    bool ok3 = (bx3 > c.xmin + c.eps) && (bx3 <= c.xmax - c.eps);
    int n = (int)ok0 + (int)ok1 + (int)ok2 + (int)ok3;
    if (n < 2) { sx = ex = c.xmin - kFar; sy = ey = c.ymin - kFar; return; }
    // first two valid candidates in the order (y=ymin, x=xmin, x=xmax, y=ymax)
    bool have = false;
    sx = sy = ex = ey = 0.f;
    if (ok0) { sx = bx0; sy = c.ymin; have = true; }
    if (ok1) { if (!have) { sx = c.xmin; sy = by1; have = true; } else { ex = c.xmin; ey = by1; return; } }
    if (ok2) { if (!have) { sx = c.xmax; sy = by2; have = true; } else { ex = c.xmax; ey = by2; return; } }
    ex = bx3; ey = c.ymax;
}

// image coordinate of sample k -> normalised grid_sample coordinate (:405-415, multiview.py:25-37,159-163)
// Evaluated with host-precomputed reciprocals (a couple of ulp from the reference's division chain,
// i.e. ~1e-6 feature px; the emitted sample_locs are exactly what the kernels sample).
__device__ __forceinline__ float img2grid_x(float v, const GeomCfg &c) { return fmaf(fmaf(v, c.inv_rds, c.off_ds), c.gsx, c.gox); }
__device__ __forceinline__ float img2grid_y(float v, const GeomCfg &c) { return fmaf(fmaf(v, c.inv_rds, c.off_ds), c.gsy, c.goy); }

// normalised grid coordinate -> source feature-pixel coordinate (ATen grid_sampler unnormalize)
// (explicit intrinsics: the union-marking code and the sampling code must produce bit-identical pixel coordinates, so
// no step may be left to the compiler's FMA contraction)
__device__ __forceinline__ float grid2pix(float g, int size, int align) {
    const float g1 = __fadd_rn(g, 1.f);
    return align ? __fmul_rn(__fmul_rn(g1, 0.5f), (float)(size - 1)) : __fmul_rn(__fmaf_rn(g1, (float)size, -1.f), 0.5f);
}
// sample at parameter t in [0,1] on the segment (sx,sy)-(ex,ey) (image coordinates) -> normalised grid coordinates
__device__ __forceinline__ float lerp_exact(float a, float b, float t) { return __fmaf_rn(__fsub_rn(b, a), t, a); }

// de_normalize (multiview.py:39-57): grid coordinate -> feature px as the reference reports corr_pos
__device__ __forceinline__ float grid2corr(float g, int size, int correct) {
    return correct ? (g + 1.f) * (float)(size - 1) * 0.5f : (g + 1.f) * (float)size * 0.5f - 0.5f;
}

struct Taps {            // bilinear footprint of one sample in the source map
    int x0, y0;          // north-west tap (may be out of bounds)
    float w[4];          // nw, ne, sw, se — already zero for out-of-bounds taps
    bool any;            // at least one tap in bounds
};

__device__ __forceinline__ Taps make_taps(float gx, float gy, int H, int W, int align) {
    Taps t;
    float ix = grid2pix(gx, W, align), iy = grid2pix(gy, H, align);
    float fx = floorf(ix), fy = floorf(iy);
    // clamp before the int conversion so far sentinels / NaN cannot overflow
    fx = fminf(fmaxf(fx, -2.f), (float)W);  fy = fminf(fmaxf(fy, -2.f), (float)H);
    float ax = ix - floorf(ix), ay = iy - floorf(iy);
    if (!(ix == ix) || !(iy == iy)) { ax = ay = 0.f; fx = fy = -2.f; }     // NaN location: all taps out
    t.x0 = (int)fx; t.y0 = (int)fy;
    bool xin0 = t.x0 >= 0 && t.x0 < W, xin1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
    bool yin0 = t.y0 >= 0 && t.y0 < H, yin1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
    t.w[0] = (xin0 && yin0) ? (1.f - ax) * (1.f - ay) : 0.f;
    t.w[1] = (xin1 && yin0) ? ax * (1.f - ay) : 0.f;
    t.w[2] = (xin0 && yin1) ? (1.f - ax) * ay : 0.f;
    t.w[3] = (xin1 && yin1) ? ax * ay : 0.f;
    t.any = (xin0 || xin1) && (yin0 || yin1);
    return t;
}


// Programmatic dependent launch (the three launches of a forward are chained with
// cudaLaunchAttributeProgrammaticStreamSerialization): a kernel's CTAs may become resident and run their prologue
// (barrier init, TMEM allocation, tensor-map prefetch) while the previous kernel of the stream drains;
// pdl_wait() returns once that kernel has completed and its writes are visible.  Nothing the previous kernels
// wrote may be read, and nothing they read may be written, before pdl_wait().
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

}  // namespace epi
