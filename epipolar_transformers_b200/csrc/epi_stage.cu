// epi_stage.cu — the ONE operand-staging launch in front of the fused attention kernel (epi_fusion_pipe.cu):
//
//   blocks [0, N)      per (ref, src) pair: fp64 pair constants (camera centre, epipole, infinite homography;
//                      /root/reference/vision/multiview.py:16-21, modeling/layers/epipolar.py:336-348) and the list of
//                      reference pixels sorted by epipolar angle (counting sort on a 12-bit angle key, ties by pixel index
//                      => deterministic).  Pixels on one epipolar line of the reference view share one epipolar line in
//                      the source view, so 32 consecutive list entries have nearly identical sets of bilinear taps.
//   remaining blocks   [N,C,H,W] fp32 (any strides) -> pixel-major bf16 (hi, lo) planes [N,H*W,C] with x ≈ hi + lo, for
//                      BOTH feature maps (reference -> planes 0,1; source -> planes 2,3 of one buffer), 64 x 64 tiles through
//                      shared memory: coalesced float4 reads along pixels, 16-byte writes along channels.
// Also zeroes the fused kernel's tile counter and error word.
#include <cuda_bf16.h>

#include "epi_kernels.cuh"

namespace epi {

namespace stg {
constexpr int NT = 256;
constexpr int NBIN = 4096;
constexpr int SMALL = 32;

__device__ __forceinline__ float angle_key(int i, int W, const GeomCfg &gc, float ex, float ey, float a0, bool parallel, float span) {
    const float px = pix2coord(i % W, gc.ds, gc.r), py = pix2coord(i / W, gc.ds, gc.r);
    float u;                                   // in [0, 1)
    if (parallel) {
        u = 0.5f + 0.5f * ((px - 0.5f * (gc.xmin + gc.xmax)) * (-ey) + (py - 0.5f * (gc.ymin + gc.ymax)) * ex) / span;
    } else {
        float ang = atan2f(py - ey, px - ex) - a0;         // relative to the image centre: the cut is behind the epipole
        if (ang < -3.14159265f) ang += 6.28318531f;
        if (ang >= 3.14159265f) ang -= 6.28318531f;
        u = (ang + 3.14159265f) * (1.f / 6.28318531f);
    }
    u = fminf(fmaxf(u, 0.f), 0.99999f);
    if (!(u == u)) u = 0.f;
    return u;
}
}  // namespace stg

#ifdef EPI_PIPE_TIMERS
__device__ unsigned long long g_stage_timers[16];
#define ST(slot) do { __syncthreads(); if (t == 0 && blockIdx.x == 0) { const long long t_ = clock64(); g_stage_timers[slot] = (unsigned long long)(t_ - st_prev); st_prev = t_; } } while (0)
extern "C" void epi_stage_timers_read(unsigned long long *out16) { cudaMemcpyFromSymbol(out16, g_stage_timers, sizeof(unsigned long long) * 16); }
#else
#define ST(slot) do { } while (0)
#endif

struct StageArgs {
    const float *ref, *src;
    int64_t ref_stride[4], src_stride[4];
    __nv_bfloat16 *planes;            // [4][N*HW][C]: ref_hi, ref_lo, src_hi, src_lo
    const float *P_ref, *P_src;       // may be null (injected locations): no order, no pair constants
    PairGeom *pair_geom;              // [N]
    uint16_t *order, *order_tmp;      // [N][HW]
    int *zero_words;                  // tile counter, error word
    int N, C, H, W;
    int do_ref, do_src, do_order;
    GeomCfg gc;
};

__global__ void __launch_bounds__(stg::NT) epi_stage_kernel(const StageArgs s) {
    using namespace stg;
    __shared__ __align__(16) float tile[64][65];          // transposition tile; the order blocks reuse it as histogram
    const int t = threadIdx.x;
    const int H = s.H, W = s.W, HW = H * W, C = s.C;
    const int nord = s.do_order ? s.N : 0;
    if (blockIdx.x == 0 && t == 0 && s.zero_words) { s.zero_words[0] = 0; s.zero_words[1] = 0; }

    if ((int)blockIdx.x < nord) {
        // ------------------------------------------------------------------------------------------------
        // pair constants + epipolar-angle order of the pair's reference pixels
        // ------------------------------------------------------------------------------------------------
        int *hist = reinterpret_cast<int *>(&tile[0][0]);              // NBIN counts -> offsets (16 KB of the 16.6 KB)
        __shared__ float s_e[4];
        __shared__ int s_warp[NT / 32];
        __shared__ int s_big[64], s_nbig;
#ifdef EPI_PIPE_TIMERS
        long long st_prev = clock64();
#endif
        const int n = blockIdx.x;
        if (t == 0) {
            const float *P1 = s.P_ref + 12 * n, *P2 = s.P_src + 12 * n;
            PairGeom g;
            pair_geom_from_krt(P1, P2, g);
            s.pair_geom[n] = g;
            // epipole of the SOURCE camera in the reference view: e1 = P_ref·[C_src; 1]
            double b[9], t2[3];
            for (int r = 0; r < 3; r++) { for (int q = 0; q < 3; q++) b[r * 3 + q] = (double)P2[r * 4 + q]; t2[r] = (double)P2[r * 4 + 3]; }
            const double c00 = b[4] * b[8] - b[5] * b[7], c01 = b[5] * b[6] - b[3] * b[8], c02 = b[3] * b[7] - b[4] * b[6];
            const double id = 1.0 / (b[0] * c00 + b[1] * c01 + b[2] * c02);
            double bi[9];
            bi[0] = c00 * id; bi[1] = (b[2] * b[7] - b[1] * b[8]) * id; bi[2] = (b[1] * b[5] - b[2] * b[4]) * id;
            bi[3] = c01 * id; bi[4] = (b[0] * b[8] - b[2] * b[6]) * id; bi[5] = (b[2] * b[3] - b[0] * b[5]) * id;
            bi[6] = c02 * id; bi[7] = (b[1] * b[6] - b[0] * b[7]) * id; bi[8] = (b[0] * b[4] - b[1] * b[3]) * id;
            double cs[3], e[3];
            for (int r = 0; r < 3; r++) cs[r] = -(bi[r * 3] * t2[0] + bi[r * 3 + 1] * t2[1] + bi[r * 3 + 2] * t2[2]);
            for (int r = 0; r < 3; r++) e[r] = (double)P1[r * 4] * cs[0] + (double)P1[r * 4 + 1] * cs[1] + (double)P1[r * 4 + 2] * cs[2] + (double)P1[r * 4 + 3];
            const double cx = 0.5 * ((double)s.gc.xmin + s.gc.xmax), cy = 0.5 * ((double)s.gc.ymin + s.gc.ymax);
            const double nrm = fabs(e[0]) + fabs(e[1]) + 1e-300;
            if (!(fabs(e[2]) > 1e-9 * nrm)) {        // epipole at infinity (or NaN): parallel lines, sort by the offset across them
                s_e[0] = (float)(e[0] / nrm); s_e[1] = (float)(e[1] / nrm); s_e[2] = 0.f; s_e[3] = 1.f;
            } else {
                const double ex = e[0] / e[2], ey = e[1] / e[2];
                s_e[0] = (float)ex; s_e[1] = (float)ey; s_e[2] = (float)atan2(cy - ey, cx - ex); s_e[3] = 0.f;
            }
            s_nbig = 0;
        }
        for (int b = t; b < NBIN; b += NT) hist[b] = 0;
        __syncthreads();
        ST(0);
        const float ex = s_e[0], ey = s_e[1], a0 = s_e[2];
        const bool parallel = s_e[3] != 0.f;
        const float span = fabsf(s.gc.xmax - s.gc.xmin) + fabsf(s.gc.ymax - s.gc.ymin) + 1.f;
        for (int i = t; i < HW; i += NT) atomicAdd(&hist[(int)(angle_key(i, W, s.gc, ex, ey, a0, parallel, span) * (float)NBIN)], 1);
        __syncthreads();
        ST(1);
        // exclusive scan over the bins: 16 consecutive bins per thread, then a block scan of the partial sums
        constexpr int PER = NBIN / NT;
        int loc[PER], sum = 0;
#pragma unroll
        for (int q = 0; q < PER; q++) { loc[q] = hist[t * PER + q]; sum += loc[q]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((t & 31) >= o) incl += v; }
        if ((t & 31) == 31) s_warp[t >> 5] = incl;
        __syncthreads();
        int base = incl - sum;
        for (int w = 0; w < (t >> 5); w++) base += s_warp[w];
        const int start0 = base;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER; q++) { hist[t * PER + q] = base; base += loc[q]; }
        __syncthreads();
        ST(2);
        // placement (arbitrary order inside a bin) into the scratch list
        uint16_t *tmp = s.order_tmp + (size_t)n * HW, *ord = s.order + (size_t)n * HW;
        for (int i = t; i < HW; i += NT) {
            const int b = (int)(angle_key(i, W, s.gc, ex, ey, a0, parallel, span) * (float)NBIN);
            tmp[atomicAdd(&hist[b], 1)] = (uint16_t)i;
        }
        __syncthreads();
        ST(3);
        // inside every bin: ascending pixel index (insertion sort; bins hold ~HW/4096 entries)
        {
            int st0 = start0;
#pragma unroll 1
            for (int q = 0; q < PER; q++) {
                const int len = loc[q];
                if (len > SMALL) {
                    const int slot = atomicAdd(&s_nbig, 1);
                    if (slot < 64) s_big[slot] = t * PER + q;
                } else if (len > 0) {
                    uint16_t v[SMALL];
                    for (int e = 0; e < len; e++) {
                        const uint16_t x = tmp[st0 + e];
                        int p = e;
                        while (p > 0 && v[p - 1] > x) { v[p] = v[p - 1]; p--; }
                        v[p] = x;
                    }
                    for (int e = 0; e < len; e++) ord[st0 + e] = v[e];
                }
                st0 += len;
            }
        }
        __syncthreads();
        ST(4);
        // degenerate cameras only: big bins are ranked by counting, the whole block per bin
        const int nbig = s_nbig < 64 ? s_nbig : 64;
        for (int bb = 0; bb < nbig; bb++) {
            const int b = s_big[bb];
            const int end = hist[b];                     // after placement: offset = end of the bin
            // start = end of the previous non-empty prefix: recompute from the neighbour (bin b-1's end), or 0
            const int beg = b == 0 ? 0 : hist[b - 1];
            for (int e = beg + t; e < end; e += NT) {
                const uint16_t x = tmp[e];
                int r = 0;
                for (int f = beg; f < end; f++) r += tmp[f] < x;
                ord[beg + r] = x;
            }
        }
        if (s_nbig > 64) {                               // pathological: more than 64 big bins — keep the scratch order
            __syncthreads();
            for (int i = t; i < HW; i += NT) ord[i] = tmp[i];
        }
        return;
    }

    // ----------------------------------------------------------------------------------------------------
    // layout staging: 64 channels x 64 pixels per block
    // ----------------------------------------------------------------------------------------------------
    const int tiles_p = (HW + 63) / 64, tiles_c = (C + 63) / 64;
    const int per_map = tiles_p * tiles_c * s.N;
    int lin = (int)blockIdx.x - nord;
    int map = 0;
    if (s.do_ref && s.do_src) { map = lin >= per_map; lin -= map * per_map; }
    else map = s.do_src ? 1 : 0;
    const int n = lin / (tiles_p * tiles_c), rem = lin % (tiles_p * tiles_c);
    const int c0 = (rem / tiles_p) * 64, p0 = (rem % tiles_p) * 64;
    const float *base = map ? s.src : s.ref;
    const int64_t *strd = map ? s.src_stride : s.ref_stride;
    const int64_t sn = strd[0], sc = strd[1], sh = strd[2], sw = strd[3];
    const float *sp = base + (int64_t)n * sn;
    const size_t plane_elems = (size_t)s.N * HW * C;
    __nv_bfloat16 *hi = s.planes + (size_t)(2 * map) * plane_elems, *lo = hi + plane_elems;
    const bool vec = (sw == 1) && (sh == W) && (HW % 4 == 0) && (sc % 4 == 0) && ((reinterpret_cast<uintptr_t>(sp) & 15) == 0);
    if (sc != 1) {
        const int q = t & 15, cy = t >> 4;                      // 16 float4 per channel row, 16 channels per pass
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int c = c0 + cy + i * 16, p = p0 + q * 4;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < C) {
                if (vec && p + 3 < HW) v[i] = __ldg(reinterpret_cast<const float4 *>(sp + c * sc + p));
                else {
                    float e[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int j = 0; j < 4; j++) if (p + j < HW) e[j] = __ldg(sp + c * sc + ((p + j) / W) * sh + ((p + j) % W) * sw);
                    v[i] = make_float4(e[0], e[1], e[2], e[3]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float *row = &tile[cy + i * 16][q * 4];
            row[0] = v[i].x; row[1] = v[i].y; row[2] = v[i].z; row[3] = v[i].w;
        }
    } else {
        const int cx = t & 63, py = t >> 6;                     // channels-last: a warp reads 32 consecutive channels
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int p = p0 + py + i * 4, c = c0 + cx;
            tile[cx][py + i * 4] = (c < C && p < HW) ? __ldg(sp + c + (p / W) * sh + (p % W) * sw) : 0.f;
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

cudaError_t launch_stage(const float *ref, const int64_t ref_stride[4], const float *src, const int64_t src_stride[4],
                         __nv_bfloat16 *planes, const float *P_ref, const float *P_src, PairGeom *pair_geom, uint16_t *order,
                         uint16_t *order_tmp, int *zero_words, int N, int C, int H, int W, const GeomCfg &gc, cudaStream_t st) {
    StageArgs s;
    s.ref = ref; s.src = src;
    for (int i = 0; i < 4; i++) { s.ref_stride[i] = ref_stride[i]; s.src_stride[i] = src_stride[i]; }
    s.planes = planes; s.P_ref = P_ref; s.P_src = P_src; s.pair_geom = pair_geom; s.order = order; s.order_tmp = order_tmp;
    s.zero_words = zero_words; s.N = N; s.C = C; s.H = H; s.W = W; s.gc = gc;
    s.do_ref = 1; s.do_src = 1; s.do_order = (P_ref && P_src && order) ? 1 : 0;
    const int tiles = ((H * W + 63) / 64) * ((C + 63) / 64) * N;
    const int grid = (s.do_order ? N : 0) + 2 * tiles;
    epi_stage_kernel<<<grid, stg::NT, 0, st>>>(s);
    return cudaGetLastError();
}

}  // namespace epi
