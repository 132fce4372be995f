// epi_umma.cuh — hand-written sm_100a plumbing for the tensor-core path: mbarrier, TMEM
// allocation, tcgen05.mma / commit / ld, shared-memory matrix descriptors and the 128-byte
// swizzle that the operand staging code must write.  Inline PTX only (no CUTLASS).
//
// Operand conventions used by the fusion kernel (bf16 operands, fp32 accumulate in TMEM):
//   "panel"  = ROWS x 64 bf16 (one 128-byte row per matrix row), rows in 8-row / 1024-byte swizzle
//              atoms (Swizzle<3,4,3>: 16-byte chunk index ^= row % 8).  A matrix with more than 64
//              columns is a sequence of panels `panel_stride` bytes apart.
//   K-major  : matrix rows are the M (or N) index, the 64 columns of a panel are consecutive K.
//   MN-major : the SAME bytes read transposed — panel rows are consecutive K, the 64 columns are
//              consecutive M.  (The gathered source-feature chunk F[d][c] is the K-major A operand of
//              S = F·Qᵀ and the MN-major A operand of O = Fᵀ·βᵀ without being re-laid out.)
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace epi {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    // suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint expires) instead of
    // spinning through the issue slots the other warps of the SM need
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- 16-byte asynchronous copies global -> shared (LDGSTS), zero-fill when !valid ---------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most n (0..3) of this thread's committed groups are still pending
__device__ __forceinline__ void cp_async_wait_pending(int n) {
    if (n <= 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
    else if (n == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else if (n == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
    else asm volatile("cp.async.wait_group 3;" ::: "memory");
}

// ---- 1-D bulk copy global -> shared (TMA engine, UBLKCP), completion on an mbarrier ----------
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- TMEM ------------------------------------------------------------------------------------
// Executed by ONE full warp; writes the allocated base address (lane<<16 | column) to *dst_smem.
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols /* power of two >= 32 */) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread `lane` of warp w receives D[32*(w%4)+lane][col..col+31].
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float *v) {
    uint32_t *r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors -------------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle (layout_type 2), descriptor version 1 (sm_100).
//   K-major : sbo = bytes between 8-row groups (1024 for dense panels), lbo ignored (1).
//   MN-major: lbo = bytes between 64-element MN groups (panel stride), sbo = bytes between 8-row K groups.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // version
    d |= (uint64_t)2 << 61;          // SWIZZLE_128B
    return d;
}

// Instruction descriptor, kind::f16, bf16 x bf16 -> fp32, dense.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4)                      // c_format  = F32
           | (1u << 7)                    // a_format  = BF16
           | (1u << 10)                   // b_format  = BF16
           | ((uint32_t)a_mn_major << 15) // a_major
           | ((uint32_t)b_mn_major << 16) // b_major
           | ((uint32_t)(N >> 3) << 17)   // n_dim
           | ((uint32_t)(M >> 4) << 24);  // m_dim
}

// D[tmem] (+)= A[smem] · B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// All previously issued MMAs of this thread arrive on the mbarrier when complete (implies fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- operand staging ----------------------------------------------------------------------------
// Byte offset of element (row, col) inside a swizzled panel set (col over all panels).
__device__ __forceinline__ uint32_t panel_offset(uint32_t row, uint32_t col, uint32_t panel_stride_bytes) {
    const uint32_t p = col >> 6, cc = col & 63;
    const uint32_t chunk = (cc >> 3) ^ (row & 7);
    return p * panel_stride_bytes + row * 128u + chunk * 16u + (cc & 7) * 2u;
}

// fp32 x -> (hi, lo) bf16 pair with x ≈ hi + lo (|x - hi - lo| <~ 2^-17 |x|)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

}  // namespace umma
}  // namespace epi
