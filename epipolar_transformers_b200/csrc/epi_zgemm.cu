// epi_zgemm.cu — z-projection epilogue on the tensor cores (tcgen05 + TMEM).
//
//   y[n,o,p] = Σ_c Wf[o,c]·x[n,c,p] + bf[o]  (+ x[n,o,p] if ZRESIDUAL)  (+ feat_ref[n,o,p] for the caller's residual)
// restates  finalout = bn(z(out)) [+ out]   /root/reference/modeling/layers/epipolar.py:249-253 (eval-mode BN folded
// into Wf, bf by epi_fold_z_bn_f32) and  ret + feat   /root/reference/modeling/backbones/resnet.py:388.
//
// One CTA per 128 pixels: D[128 px, C out] = X[128 px, C]·Wfᵀ with X supplied by the fusion kernel as bf16
// (hi, lo) planes [N·HW, C] (K-major rows) and Wf split to (hi, lo) while it is staged.  Three MMAs per
// product (hi·hi + hi·lo + lo·hi), fp32 accumulation in TMEM (M=128, N=C<=256), K streamed in 64-channel
// panels through a double-buffered shared-memory ring — the X panels arrive by TMA (cp.async.bulk.tensor.2d with the
// 128-byte swizzle the UMMA descriptors expect, completion on an mbarrier); the epilogue adds bias/residuals and writes NCHW
// (a warp's lanes are 32 consecutive pixels, so every store instruction is one 128-byte line per channel).
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstring>

#include "epi_kernels.cuh"
#include "epi_umma.cuh"

namespace epi {
using namespace umma;

namespace zg {
constexpr int NT = 256;
constexpr uint32_t A_PLANE = 16384;                 // 128 rows x 128 B
constexpr uint32_t B_PLANE = 32768;                 // 256 rows x 128 B
constexpr uint32_t STAGE = 2 * A_PLANE + 2 * B_PLANE;   // 96 KB
constexpr uint32_t SMEM_ALLOC = 2 * STAGE + 1024 + 64;

// 2-D tiled TMA load of a [128 rows x 64 bf16] box into a swizzled panel, completion counted on `bar`
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     umma::smem_u32(smem_dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(umma::smem_u32(bar))
                 : "memory");
}
}  // namespace zg

__global__ void __launch_bounds__(zg::NT, 1) epi_zgemm_kernel(const ZGemmArgs z, const __grid_constant__ CUtensorMap tm_hi,
                                                              const __grid_constant__ CUtensorMap tm_lo, const int use_tma) {
    using namespace zg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * STAGE);       // [0,1] MMAs done with stage, [2] all, [3] TMA landed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 2 * STAGE + 32);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = z.C, HW = z.HW, W = z.W;
    const int tiles = (HW + 127) / 128;
    const int n = blockIdx.x / tiles, p0 = (blockIdx.x % tiles) * 128;
    const int nq = (C + 63) / 64;                      // K panels of 64 channels
    const __nv_bfloat16 *xh = z.x_hi + (size_t)n * HW * C, *xl = z.x_lo + (size_t)n * HW * C;

    if (warp == 0) tmem_alloc(tmem_slot, 256);
    if (tid == 32) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], 1); mbar_fence_init(); }
    if (tid == 0 && use_tma) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_lo) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    for (int q = 0; q < nq; q++) {
        const uint32_t buf = q & 1;
        uint8_t *st = smem + buf * STAGE;
        if (q >= 2) { for (uint32_t it = 0; !mbar_try_wait(&bars[buf], ((q >> 1) + 1) & 1); ++it) if (it > (1u << 26)) __trap(); }
        // A: 128 pixel rows x 64 channels of both planes
        if (use_tma) {
            if (tid == 0) {                                     // one thread arms the barrier and issues both boxes
                mbar_arrive_expect_tx(&bars[3], 2 * A_PLANE);
                tma_load_2d(st, &tm_hi, q * 64, n * HW + p0, &bars[3]);
                tma_load_2d(st + A_PLANE, &tm_lo, q * 64, n * HW + p0, &bars[3]);
            }
        } else {                                                // 16-byte chunks, 8 lanes per row
            const int j = tid & 7;
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int plane = it >> 2, r = (it & 3) * 32 + (tid >> 3), p = p0 + r, c0 = q * 64 + j * 8;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (p < HW && c0 < C) v = __ldg(reinterpret_cast<const uint4 *>((plane ? xl : xh) + (size_t)p * C + c0));
                *reinterpret_cast<uint4 *>(st + plane * A_PLANE + r * 128u + ((j ^ (r & 7)) << 4)) = v;
            }
        }
        // B: Wf rows o (out channels) x 64 input channels, split to (hi, lo) on the fly
        {
            const int j = tid & 7;
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int o = it * 32 + (tid >> 3), c0 = q * 64 + j * 8;
                float f[8];
#pragma unroll
                for (int u = 0; u < 8; u++) f[u] = 0.f;
                if (o < C && c0 < C) {
                    const float4 a = __ldg(reinterpret_cast<const float4 *>(z.Wf + (size_t)o * C + c0));
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(z.Wf + (size_t)o * C + c0 + 4));
                    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
                }
                uint32_t h[4], l[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const __nv_bfloat162 hv = __floats2bfloat162_rn(f[2 * u], f[2 * u + 1]);
                    const float2 hf = __bfloat1622float2(hv);
                    const __nv_bfloat162 lv = __floats2bfloat162_rn(f[2 * u] - hf.x, f[2 * u + 1] - hf.y);
                    h[u] = *reinterpret_cast<const uint32_t *>(&hv);
                    l[u] = *reinterpret_cast<const uint32_t *>(&lv);
                }
                const uint32_t off = 2 * A_PLANE + o * 128u + ((j ^ (o & 7)) << 4);
                *reinterpret_cast<uint4 *>(st + off) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4 *>(st + B_PLANE + off) = make_uint4(l[0], l[1], l[2], l[3]);
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            if (use_tma) { for (uint32_t it = 0; !mbar_try_wait(&bars[3], q & 1); ++it) if (it > (1u << 26)) __trap(); }
            const uint32_t idesc = make_idesc_bf16(128, z.Npad, 0, 0);
            const uint32_t sa = smem_u32(st), sb = sa + 2 * A_PLANE;
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                const uint64_t a_hi = make_smem_desc(sa + ks * 32, 16, 1024), a_lo = make_smem_desc(sa + A_PLANE + ks * 32, 16, 1024);
                const uint64_t b_hi = make_smem_desc(sb + ks * 32, 16, 1024), b_lo = make_smem_desc(sb + B_PLANE + ks * 32, 16, 1024);
                mma_bf16(tmem, a_hi, b_hi, idesc, (q | ks) ? 1u : 0u);
                mma_bf16(tmem, a_hi, b_lo, idesc, 1u);
                mma_bf16(tmem, a_lo, b_hi, idesc, 1u);
            }
            mma_commit(&bars[buf]);
        }
    }
    if (tid == 0) mma_commit(&bars[2]);
    for (uint32_t it = 0; !mbar_try_wait(&bars[2], 0); ++it) if (it > (1u << 26)) __trap();
    tc_fence_after();

    // epilogue: thread <-> pixel (TMEM lane), 32 output channels at a time
    {
        const int r = (warp & 3) * 32 + lane, p = p0 + r;
        const bool ok = p < HW;
        const int py = ok ? p / W : 0, px = ok ? p % W : 0;
        float *yb = z.y + (int64_t)n * z.y_stride[0] + (int64_t)py * z.y_stride[2] + (int64_t)px * z.y_stride[3];
        const float *rb = z.ref ? z.ref + (int64_t)n * z.ref_stride[0] + (int64_t)py * z.ref_stride[2] + (int64_t)px * z.ref_stride[3] : nullptr;
        for (int cb = (warp >> 2) * 32; cb < C; cb += 64) {
            float v[32];
            tmem_ld_32x32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + cb, v);
            tmem_ld_wait();
            if (ok) {
                uint4 xh4[4], xl4[4];
                if (z.z_residual) {
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        xh4[u] = (cb + u * 8 < C) ? __ldg(reinterpret_cast<const uint4 *>(xh + (size_t)p * C + cb + u * 8)) : make_uint4(0, 0, 0, 0);
                        xl4[u] = (cb + u * 8 < C) ? __ldg(reinterpret_cast<const uint4 *>(xl + (size_t)p * C + cb + u * 8)) : make_uint4(0, 0, 0, 0);
                    }
                }
#pragma unroll
                for (int jj = 0; jj < 32; jj++) {
                    const int o = cb + jj;
                    if (o < C) {
                        float y = v[jj] + __ldg(z.bf + o);
                        if (z.z_residual) {
                            const uint32_t wh = reinterpret_cast<const uint32_t *>(xh4)[jj >> 1], wl = reinterpret_cast<const uint32_t *>(xl4)[jj >> 1];
                            const uint32_t bh = (jj & 1) ? (wh & 0xffff0000u) : (wh << 16), bl = (jj & 1) ? (wl & 0xffff0000u) : (wl << 16);
                            y += __uint_as_float(bh) + __uint_as_float(bl);
                        }
                        if (z.add_ref && rb) y += __ldg(rb + (int64_t)o * z.ref_stride[1]);
                        yb[(int64_t)o * z.y_stride[1]] = y;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

bool zgemm_supported(int C) { return C % 16 == 0 && C >= 16 && C <= 256; }

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// [rows = N*HW, cols = C] bf16 plane, box = 64 columns x 128 rows, 128-byte swizzle (what the UMMA K-major descriptor reads)
bool make_plane_map(CUtensorMap *m, const __nv_bfloat16 *base, int rows, int C) {
    EncodeTiledFn fn = encode_fn();
    if (!fn || C % 8 != 0) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    const cuuint32_t box[2] = {64u, 128u};
    const cuuint32_t estr[2] = {1u, 1u};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16 *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

cudaError_t launch_zgemm(const ZGemmArgs &z, cudaStream_t st) {
    const int tiles = (z.HW + 127) / 128;
    cudaError_t e = cudaFuncSetAttribute(epi_zgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zg::SMEM_ALLOC);
    if (e != cudaSuccess) return e;
    CUtensorMap tm_hi, tm_lo;
    memset(&tm_hi, 0, sizeof(tm_hi)); memset(&tm_lo, 0, sizeof(tm_lo));
    // TMA needs whole 64-column boxes inside the row pitch; otherwise the LDG path stages A
    const int use_tma = (z.C % 64 == 0) && make_plane_map(&tm_hi, z.x_hi, z.N * z.HW, z.C) && make_plane_map(&tm_lo, z.x_lo, z.N * z.HW, z.C);
    epi_zgemm_kernel<<<z.N * tiles, zg::NT, zg::SMEM_ALLOC, st>>>(z, tm_hi, tm_lo, use_tma);
    return cudaGetLastError();
}

}  // namespace epi
