// epi_zgemm.cu — z-projection epilogue on the tensor cores (tcgen05 + TMEM).
//
//   y[n,o,p] = Σ_c Wf[o,c]·x[n,c,p] + bf[o]  (+ x[n,o,p] if ZRESIDUAL)  (+ feat_ref[n,o,p] for the caller's residual)
// restates  finalout = bn(z(out)) [+ out]   /root/reference/modeling/layers/epipolar.py:249-253 (eval-mode BN folded
// into Wf, bf by epi_fold_z_bn_f32) and  ret + feat   /root/reference/modeling/backbones/resnet.py:388.
//
// One CTA per 128 pixels and block of up to 256 output channels (blockIdx.y; two blocks when 256 < C <= 512):
// D[128 px, C out] = X[128 px, C]·Wfᵀ with X supplied by the fusion kernel as bf16
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
constexpr int NT = 512;
constexpr uint32_t A_PLANE = 16384;                 // 128 rows x 128 B
constexpr uint32_t B_PLANE = 32768;                 // 256 rows x 128 B
constexpr uint32_t STAGE = 2 * A_PLANE + 2 * B_PLANE;   // 96 KB
constexpr uint32_t SMEM_ALLOC = 2 * STAGE + 1024 + 128;

// 2-D tiled TMA load of a [128 rows x 64 bf16] box into a swizzled panel, completion counted on `bar`
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     umma::smem_u32(smem_dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(umma::smem_u32(bar))
                 : "memory");
}
}  // namespace zg

__global__ void __launch_bounds__(zg::NT, 1) epi_zgemm_kernel(const ZGemmArgs z, const __grid_constant__ CUtensorMap tm_hi,
                                                              const __grid_constant__ CUtensorMap tm_lo,
                                                              const __grid_constant__ CUtensorMap tw_hi,
                                                              const __grid_constant__ CUtensorMap tw_lo) {
    using namespace zg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * STAGE);       // [0,1] stage landed (TMA), [2,3] MMAs done with stage, [4] all MMAs done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 2 * STAGE + 64);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = z.C, HW = z.HW, W = z.W;
    const int tiles = (HW + 127) / 128;
    const int n = blockIdx.x / tiles, p0 = (blockIdx.x % tiles) * 128;
    const int nq = (C + 63) / 64;                      // K panels of 64 channels
    const int oc0 = (int)blockIdx.y * 256;            // this CTA's block of output channels
    const int CO = min(256, C - oc0);
    const __nv_bfloat16 *xh = z.x_hi + (size_t)n * HW * C, *xl = z.x_lo + (size_t)n * HW * C;

    pdl_launch_dependents();
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    if (tid == 32) { for (int i = 0; i < 5; i++) mbar_init(&bars[i], 1); mbar_fence_init(); }
    if (tid == 64) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tw_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tw_lo) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();                                        // the fused kernel's feature planes

    // One thread drives the whole main loop: TMA loads of the (A, W) panels into a 2-stage ring, tcgen05.mma issue, commits.
    // A: 128 pixel rows x 64 channels of the (hi, lo) planes of x;  B: C output rows x 64 input channels of the (hi, lo) planes of Wf.
    if (tid == 0) {
        auto load = [&](int q) {
            uint8_t *st = smem + (q & 1) * STAGE;
            mbar_arrive_expect_tx(&bars[q & 1], 2 * A_PLANE + 2 * (uint32_t)(C < 256 ? C : 256) * 128u);   // W box = min(C, 256) rows x 128 B
            tma_load_2d(st, &tm_hi, q * 64, n * HW + p0, &bars[q & 1]);
            tma_load_2d(st + A_PLANE, &tm_lo, q * 64, n * HW + p0, &bars[q & 1]);
            tma_load_2d(st + 2 * A_PLANE, &tw_hi, q * 64, oc0, &bars[q & 1]);
            tma_load_2d(st + 2 * A_PLANE + B_PLANE, &tw_lo, q * 64, oc0, &bars[q & 1]);
        };
        load(0);
        if (nq > 1) load(1);
        const uint32_t idesc = make_idesc_bf16(128, (uint32_t)((CO + 15) & ~15), 0, 0);
        for (int q = 0; q < nq; q++) {
            const uint32_t buf = q & 1;
            for (uint32_t it = 0; !mbar_try_wait(&bars[buf], (q >> 1) & 1); ++it) if (it > (1u << 24)) __trap();
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + buf * STAGE), sb = sa + 2 * A_PLANE;
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                const uint64_t a_hi = make_smem_desc(sa + ks * 32, 16, 1024), a_lo = make_smem_desc(sa + A_PLANE + ks * 32, 16, 1024);
                const uint64_t b_hi = make_smem_desc(sb + ks * 32, 16, 1024), b_lo = make_smem_desc(sb + B_PLANE + ks * 32, 16, 1024);
                mma_bf16(tmem, a_hi, b_hi, idesc, (q | ks) ? 1u : 0u);
                mma_bf16(tmem, a_hi, b_lo, idesc, 1u);
                mma_bf16(tmem, a_lo, b_hi, idesc, 1u);
            }
            mma_commit(&bars[2 + buf]);
            if (q + 2 < nq) {
                for (uint32_t it = 0; !mbar_try_wait(&bars[2 + buf], (q >> 1) & 1); ++it) if (it > (1u << 24)) __trap();
                load(q + 2);
            }
        }
        mma_commit(&bars[4]);
    }
    for (uint32_t it = 0; !mbar_try_wait(&bars[4], 0); ++it) if (it > (1u << 24)) __trap();
    tc_fence_after();

    // ---- epilogue ------------------------------------------------------------------------------------------------
    // phase 1 (thread <-> pixel = TMEM lane, 16 warps = 4 lane quadrants x 4 channel groups): accumulator -> shared tile
    //   [channel][128 pixels] (the operand stages are free now).  The ZRESIDUAL needs no pass: the staged weight is Wf + I;
    // phase 2 (warp <-> channel row, lane <-> 4 consecutive pixels): + bias + caller residual, 512-byte row segments of the NCHW output
    //   per warp instruction.  Falls back to per-element addressing for strides that are not pixel-contiguous.
    float *otile = reinterpret_cast<float *>(smem);            // [C][132] fp32 <= 256 * 132 * 4 = 135168 B of the 196608 B stage area
    constexpr int OT = 132;
    {
        const int r = (warp & 3) * 32 + lane;
        for (int cb = (warp >> 2) * 32; cb < CO; cb += 128) {
            float v[32];
            tmem_ld_32x32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + cb, v);
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 32; jj++)
                if (cb + jj < CO) otile[(cb + jj) * OT + r] = v[jj];
        }
    }
    tc_fence_before();
    __syncthreads();
    {
        const bool addr = z.ref && z.add_ref;
        const bool vec = (z.y_stride[3] == 1) && (z.y_stride[2] == W) && (HW % 4 == 0) && (z.y_stride[1] % 4 == 0) && (z.y_stride[0] % 4 == 0) &&
                         ((reinterpret_cast<uintptr_t>(z.y) & 15) == 0) &&
                         (!addr || ((z.ref_stride[3] == 1) && (z.ref_stride[2] == W) && (z.ref_stride[1] % 4 == 0) && (z.ref_stride[0] % 4 == 0) &&
                                    ((reinterpret_cast<uintptr_t>(z.ref) & 15) == 0)));
        const int pp = lane * 4, p = p0 + pp;
        for (int ol = warp; ol < CO; ol += NT / 32) {
            const int o = oc0 + ol;
            const float4 t = *reinterpret_cast<const float4 *>(otile + ol * OT + pp);
            const float b = __ldg(z.bf + o);
            float y[4] = {t.x + b, t.y + b, t.z + b, t.w + b};
            if (vec && p + 3 < HW) {
                if (addr) {
                    const float4 r4 = __ldg(reinterpret_cast<const float4 *>(z.ref + (int64_t)n * z.ref_stride[0] + (int64_t)o * z.ref_stride[1] + p));
                    y[0] += r4.x; y[1] += r4.y; y[2] += r4.z; y[3] += r4.w;
                }
                *reinterpret_cast<float4 *>(z.y + (int64_t)n * z.y_stride[0] + (int64_t)o * z.y_stride[1] + p) = make_float4(y[0], y[1], y[2], y[3]);
            } else {
                for (int e = 0; e < 4 && p + e < HW; e++) {
                    const int py = (p + e) / W, px = (p + e) % W;
                    float val = y[e];
                    if (addr) val += __ldg(z.ref + (int64_t)n * z.ref_stride[0] + (int64_t)o * z.ref_stride[1] + (int64_t)py * z.ref_stride[2] + (int64_t)px * z.ref_stride[3]);
                    z.y[(int64_t)n * z.y_stride[0] + (int64_t)o * z.y_stride[1] + (int64_t)py * z.y_stride[2] + (int64_t)px * z.y_stride[3]] = val;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

bool zgemm_supported(int C) { return C % 64 == 0 && C >= 64 && C <= 512; }   // whole 64-channel TMA panels; two output blocks above 256

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
// [rows, cols = C] bf16 plane, box = 64 columns x box_rows rows, 128-byte swizzle (what the UMMA K-major descriptor reads)
bool make_plane_map(CUtensorMap *m, const __nv_bfloat16 *base, int rows, int C, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn || C % 8 != 0) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16 *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

cudaError_t launch_zgemm(const ZGemmArgs &z, cudaStream_t st) {
    const int tiles = (z.HW + 127) / 128;
    static thread_local bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(epi_zgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zg::SMEM_ALLOC);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    // tensor maps are a pure function of (pointers, shape): keep the last set per host thread
    struct MapCache { const void *xh, *wh; int rows, C; CUtensorMap m[4]; };
    static thread_local MapCache mc = {nullptr, nullptr, 0, 0, {}};
    if (mc.xh != z.x_hi || mc.wh != z.w_hi || mc.rows != z.N * z.HW || mc.C != z.C) {
        const int wrows = z.C < 256 ? z.C : 256;
        if (!make_plane_map(&mc.m[0], z.x_hi, z.N * z.HW, z.C, 128) || !make_plane_map(&mc.m[1], z.x_lo, z.N * z.HW, z.C, 128) ||
            !make_plane_map(&mc.m[2], z.w_hi, z.C, z.C, wrows) || !make_plane_map(&mc.m[3], z.w_lo, z.C, z.C, wrows))
            return cudaErrorInvalidValue;
        mc.xh = z.x_hi; mc.wh = z.w_hi; mc.rows = z.N * z.HW; mc.C = z.C;
    }
    return launch_pdl(epi_zgemm_kernel, dim3((unsigned)(z.N * tiles), (unsigned)((z.C + 255) / 256)), dim3(zg::NT), (size_t)zg::SMEM_ALLOC, st, z, mc.m[0], mc.m[1], mc.m[2], mc.m[3]);
}

}  // namespace epi
