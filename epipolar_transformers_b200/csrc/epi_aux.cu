// epi_aux.cu — layout staging, parameter folding, the z/BN epilogue and the geometry-only kernel.
#include "epi_kernels.cuh"

namespace epi {

// ------------------------------------------------------------------------------------------
// [N,C,H,W] (any strides) -> [N,H,W,C] contiguous.  32x32 shared-memory transpose per item:
// reads are coalesced along the pixel axis when stride[3]==1 (NCHW), writes along channels.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float *__restrict__ src, int64_t sn, int64_t sc,
                                                           int64_t sh, int64_t sw, float *__restrict__ dst,
                                                           int C, int H, int W) {
    __shared__ float tile[32][33];
    const int HW = H * W;
    const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    const float *s = src + (int64_t)n * sn;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int c = c0 + ty + i * 8, p = p0 + tx;
        tile[ty + i * 8][tx] = (c < C && p < HW) ? __ldg(s + c * sc + (p / W) * sh + (p % W) * sw) : 0.f;
    }
    __syncthreads();
    float *d = dst + (size_t)n * HW * C;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int p = p0 + ty + i * 8, c = c0 + tx;
        if (c < C && p < HW) d[(size_t)p * C + c] = tile[tx][ty + i * 8];
    }
}

cudaError_t launch_nchw_to_nhwc(const float *src, const int64_t stride[4], float *dst, int N, int C, int H, int W,
                                cudaStream_t st) {
    dim3 grid((H * W + 31) / 32, (C + 31) / 32, N);
    nchw_to_nhwc_kernel<<<grid, 256, 0, st>>>(src, stride[0], stride[1], stride[2], stride[3], dst, C, H, W);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// [N,C,H,W] fp32 (any strides) -> two bf16 planes [N,H,W,C] with src ≈ hi + lo (operands of the
// tensor-core kernel; |src - hi - lo| <~ 2^-17 |src|).  Same 32x32 transpose tile as above.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_planes_kernel(const float *__restrict__ src, int64_t sn, int64_t sc, int64_t sh,
                                                           int64_t sw, __nv_bfloat16 *__restrict__ hi,
                                                           __nv_bfloat16 *__restrict__ lo, int C, int H, int W, int *zero_me) {
    // tile: 64 channels x 64 pixels.  NCHW sources are read as float4 (a warp covers two 256-byte runs, which also
    // keeps NVLink requests large when `src` is a peer-mapped map of another GPU); planes are written as 16-byte
    // chunks of 8 channels.
    __shared__ float tile[64][65];
    if (zero_me && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) *zero_me = 0;   // tile scheduler counter
    const int HW = H * W;
    const int n = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
    const int t = threadIdx.x;
    const float *s = src + (int64_t)n * sn;
    const bool vec = (sw == 1) && (sh == W) && (HW % 4 == 0) && (sc % 4 == 0) && ((reinterpret_cast<uintptr_t>(s) & 15) == 0);
    if (sc != 1) {
        const int q = t & 15, cy = t >> 4;                      // 16 float4 per channel row, 16 channels per pass
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int c = c0 + cy + i * 16, p = p0 + q * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < C) {
                if (vec && p + 3 < HW) v = __ldg(reinterpret_cast<const float4 *>(s + c * sc + p));
                else {
                    float e[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int j = 0; j < 4; j++) if (p + j < HW) e[j] = __ldg(s + c * sc + ((p + j) / W) * sh + ((p + j) % W) * sw);
                    v = make_float4(e[0], e[1], e[2], e[3]);
                }
            }
            float *row = &tile[cy + i * 16][q * 4];
            row[0] = v.x; row[1] = v.y; row[2] = v.z; row[3] = v.w;
        }
    } else {
        const int cx = t & 63, py = t >> 6;                     // channels-last: a warp reads 32 consecutive channels
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int p = p0 + py + i * 4, c = c0 + cx;
            tile[cx][py + i * 4] = (c < C && p < HW) ? __ldg(s + c + (p / W) * sh + (p % W) * sw) : 0.f;
        }
    }
    __syncthreads();
    const int cg = t & 7, pl = t >> 3;                          // 8 channel groups x 32 pixels per pass
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int pp = pl + i * 32, p = p0 + pp, c = c0 + cg * 8;
        if (p < HW && c < C) {                                  // C % 8 == 0 on this path
            uint32_t h[4], l[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float f0 = tile[cg * 8 + 2 * u][pp], f1 = tile[cg * 8 + 2 * u + 1][pp];
                const __nv_bfloat162 hv = __floats2bfloat162_rn(f0, f1);
                const float2 hf = __bfloat1622float2(hv);
                const __nv_bfloat162 lv = __floats2bfloat162_rn(f0 - hf.x, f1 - hf.y);
                h[u] = *reinterpret_cast<const uint32_t *>(&hv);
                l[u] = *reinterpret_cast<const uint32_t *>(&lv);
            }
            const size_t o = ((size_t)n * HW + p) * C + c;
            *reinterpret_cast<uint4 *>(hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4 *>(lo + o) = make_uint4(l[0], l[1], l[2], l[3]);
        }
    }
}

cudaError_t launch_split_planes(const float *src, const int64_t stride[4], __nv_bfloat16 *hi, __nv_bfloat16 *lo, int N, int C,
                                int H, int W, int *zero_me, cudaStream_t st) {
    dim3 grid((H * W + 63) / 64, (C + 63) / 64, N);
    split_planes_kernel<<<grid, 256, 0, st>>>(src, stride[0], stride[1], stride[2], stride[3], hi, lo, C, H, W, zero_me);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Pixel-major fp32 plane [N,H*W,C] (what the fused kernel writes with full 128-byte lines) -> the caller's
// [N,C,H,W] tensor (any strides), optionally adding the caller's residual feat_ref (resnet.py:388).  64 x 64
// tiles through shared memory: float4 reads along channels, float4 writes along pixels.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unstage_kernel(const float *__restrict__ pm, const float *__restrict__ ref, int64_t rn,
                                                      int64_t rc, int64_t rh, int64_t rw, float *__restrict__ out, int64_t on,
                                                      int64_t oc, int64_t oh, int64_t ow, int C, int H, int W) {
    __shared__ float tile[64][65];                              // [channel][pixel]
    const int HW = H * W, t = threadIdx.x;
    const int n = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
    {
        const int q = t & 15, pl = t >> 4;                      // 16 float4 per pixel row (64 channels), 16 pixels per pass
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int p = p0 + pl + i * 16, c = c0 + q * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < HW && c < C) v = __ldg(reinterpret_cast<const float4 *>(pm + ((size_t)n * HW + p) * C + c));   // C % 4 == 0
            tile[q * 4 + 0][pl + i * 16] = v.x; tile[q * 4 + 1][pl + i * 16] = v.y;
            tile[q * 4 + 2][pl + i * 16] = v.z; tile[q * 4 + 3][pl + i * 16] = v.w;
        }
    }
    __syncthreads();
    const bool vec_o = (ow == 1) && (oh == W) && (HW % 4 == 0) && (oc % 4 == 0) && (on % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const bool vec_r = !ref || ((rw == 1) && (rh == W) && (rc % 4 == 0) && (rn % 4 == 0) && ((reinterpret_cast<uintptr_t>(ref) & 15) == 0));
    const int q = t & 15, cy = t >> 4;                          // 16 float4 per channel row, 16 channels per pass
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int c = c0 + cy + i * 16, p = p0 + q * 4;
        if (c >= C || p >= HW) continue;
        float v[4] = {tile[cy + i * 16][q * 4], tile[cy + i * 16][q * 4 + 1], tile[cy + i * 16][q * 4 + 2], tile[cy + i * 16][q * 4 + 3]};
        if (vec_o && vec_r && p + 3 < HW) {
            if (ref) {
                const float4 r4 = __ldg(reinterpret_cast<const float4 *>(ref + (int64_t)n * rn + (int64_t)c * rc + p));
                v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
            }
            *reinterpret_cast<float4 *>(out + (int64_t)n * on + (int64_t)c * oc + p) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            for (int j = 0; j < 4 && p + j < HW; j++) {
                const int y = (p + j) / W, x = (p + j) % W;
                float o = v[j];
                if (ref) o += __ldg(ref + (int64_t)n * rn + (int64_t)c * rc + (int64_t)y * rh + (int64_t)x * rw);
                out[(int64_t)n * on + (int64_t)c * oc + (int64_t)y * oh + (int64_t)x * ow] = o;
            }
        }
    }
}

cudaError_t launch_unstage(const float *pm, const float *ref, const int64_t ref_stride[4], float *out, const int64_t out_stride[4],
                           int N, int C, int H, int W, cudaStream_t st) {
    dim3 grid((H * W + 63) / 64, (C + 63) / 64, N);
    unstage_kernel<<<grid, 256, 0, st>>>(pm, ref, ref ? ref_stride[0] : 0, ref ? ref_stride[1] : 0, ref ? ref_stride[2] : 0,
                                         ref ? ref_stride[3] : 0, out, out_stride[0], out_stride[1], out_stride[2], out_stride[3], C, H, W);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Fold conv1x1 z + eval BN (epipolar.py:250-251, BN.py:79 with training=False) into Wf, bf.
// ------------------------------------------------------------------------------------------
__global__ void fold_z_bn_kernel(const float *__restrict__ zw, const float *__restrict__ zb,
                                 const float *__restrict__ g, const float *__restrict__ b,
                                 const float *__restrict__ mean, const float *__restrict__ var, float eps, int C,
                                 float *__restrict__ wf, float *__restrict__ bf) {
    const int o = blockIdx.x;
    const float s = g[o] / sqrtf(var[o] + eps);
    for (int c = threadIdx.x; c < C; c += blockDim.x) wf[(size_t)o * C + c] = s * zw[(size_t)o * C + c];
    if (threadIdx.x == 0) bf[o] = s * ((zb ? zb[o] : 0.f) - mean[o]) + b[o];
}

cudaError_t launch_fold_z_bn(const float *zw, const float *zb, const float *g, const float *b, const float *mean,
                             const float *var, float eps, int C, float *wf, float *bf, cudaStream_t st) {
    fold_z_bn_kernel<<<C, 128, 0, st>>>(zw, zb, g, b, mean, var, eps, C, wf, bf);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// z epilogue: per item Y[C x HW] = Wf[C x C] · X[C x HW] + bf (+X) (+ref).  fp32 CUDA-core
// SGEMM, 64x64 tile, 4x4 per thread.  X is the library's own contiguous pre-z buffer.
// ------------------------------------------------------------------------------------------
constexpr int ZT = 64, ZK = 16;

__global__ void __launch_bounds__(256) z_epilogue_kernel(const ZArgs z) {
    __shared__ float Ws[ZK][ZT + 1];     // [k][o]
    __shared__ float Xs[ZK][ZT + 1];     // [k][p]
    const int n = blockIdx.z, o0 = blockIdx.y * ZT, p0 = blockIdx.x * ZT;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;       // tx -> p, ty -> o
    const int C = z.C, HW = z.HW, W = z.W;
    const float *X = z.x + (int64_t)n * z.x_stride[0];
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < C; k0 += ZK) {
        for (int idx = tid; idx < ZK * ZT; idx += 256) {
            int kk = idx & (ZK - 1), oo = idx / ZK;                   // W rows are contiguous in c
            int o = o0 + oo, c = k0 + kk;
            Ws[kk][oo] = (o < C && c < C) ? __ldg(z.Wf + (size_t)o * C + c) : 0.f;
            int pp = idx & (ZT - 1), kx = idx / ZT;
            int p = p0 + pp, cx = k0 + kx;
            Xs[kx][pp] = (p < HW && cx < C) ? __ldg(X + cx * z.x_stride[1] + (p / W) * z.x_stride[2] + (p % W) * z.x_stride[3]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < ZK; kk++) {
            float wv[4], xv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { wv[i] = Ws[kk][ty * 4 + i]; xv[i] = Xs[kk][tx + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *Y = z.y + (int64_t)n * z.y_stride[0];
    const float *R = z.ref ? z.ref + (int64_t)n * z.ref_stride[0] : nullptr;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int o = o0 + ty * 4 + i;
        if (o >= C) continue;
        float b = __ldg(z.bf + o);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int p = p0 + tx + 16 * j;
            if (p >= HW) continue;
            int yy = p / W, xx = p % W;
            float v = acc[i][j] + b;
            if (z.z_residual) v += __ldg(X + o * z.x_stride[1] + yy * z.x_stride[2] + xx * z.x_stride[3]);
            if (z.add_ref && R) v += __ldg(R + o * z.ref_stride[1] + yy * z.ref_stride[2] + xx * z.ref_stride[3]);
            Y[o * z.y_stride[1] + yy * z.y_stride[2] + xx * z.y_stride[3]] = v;
        }
    }
}

cudaError_t launch_z_epilogue(const ZArgs &z, cudaStream_t st) {
    dim3 grid((z.HW + ZT - 1) / ZT, (z.C + ZT - 1) / ZT, z.N);
    z_epilogue_kernel<<<grid, 256, 0, st>>>(z);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// geometry only: sample locations [K,N,H,W,2] (grid2sample_locs, epipolar.py:323-418)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_locs_kernel(const float *__restrict__ P_ref, const float *__restrict__ P_src,
                                                          float *__restrict__ locs, int N, const GeomCfg gc) {
    __shared__ PairGeom sg;
    const int n = blockIdx.y, HW = gc.H * gc.W;
    if (threadIdx.x == 0) pair_geom_from_krt(P_ref + 12 * n, P_src + 12 * n, sg);
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    float sx, sy, ex, ey;
    line_endpoints(sg, gc, pix2coord(p % gc.W, gc.ds, gc.r), pix2coord(p / gc.W, gc.ds, gc.r), sx, sy, ex, ey);
    for (int k = 0; k < gc.K; k++) {
        float t = (float)k / (float)(gc.K - 1);
        reinterpret_cast<float2 *>(locs)[((size_t)k * N + n) * HW + p] =
            make_float2(img2grid_x(sx + (ex - sx) * t, gc), img2grid_y(sy + (ey - sy) * t, gc));
    }
}

cudaError_t launch_sample_locs(const float *P_ref, const float *P_src, float *locs, int N, const GeomCfg &gc,
                               cudaStream_t st) {
    dim3 grid((gc.H * gc.W + 255) / 256, N);
    sample_locs_kernel<<<grid, 256, 0, st>>>(P_ref, P_src, locs, N, gc);
    return cudaGetLastError();
}

}  // namespace epi
