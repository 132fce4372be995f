// gather4_probe.cu — stand-alone probe (no torch): does cp.async.bulk.tensor.2d...tile::gather4 land four arbitrary rows
// of a [rows][C] bf16 plane in a 128-byte-swizzled shared-memory panel the way the UMMA descriptors expect, and how
// fast can one SM / the whole chip pull L2-resident rows with it?   nvcc -arch=sm_100a -o gather4_probe gather4_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void gather4(void *dst, const CUtensorMap *tm, int col, int r0, int r1, int r2, int r3, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}

// ---- correctness: gather 128 rows x 64 columns (one panel), dump shared memory raw -------------------------------
__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, const int *rows, int col0, uint4 *dump) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 16384);
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) mbar_expect_tx(bar, 16384);
    __syncthreads();
    if (threadIdx.x < 32) {
        const int g = threadIdx.x;
        gather4(smem + g * 512, &tm, col0, rows[4 * g], rows[4 * g + 1], rows[4 * g + 2], rows[4 * g + 3], bar);
    }
    for (uint32_t it = 0; !mbar_try_wait(bar, 0); ++it) if (it > (1u << 24)) { if (threadIdx.x == 0) printf("TIMEOUT\n"); return; }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) dump[i] = reinterpret_cast<uint4 *>(smem)[i];
}

// ---- throughput: every CTA streams `iters` stages of 32 KB (64 gather4) through a ring of `depth` buffers --------
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, const int *rows, int nrows_list,
                                                      int iters, int depth, unsigned long long *cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 6 * 32768);
    if (threadIdx.x == 0) { for (int i = 0; i < 6; i++) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long t0 = clock64();
    if (warp == 0) {
        // producer warp; consumer = same warp waiting on stage (it - depth) before re-arming (no MMA here: pure TMA rate)
        for (int it = 0; it < iters; it++) {
            const int s = it % depth;
            if (it >= depth) { const uint32_t par = ((it / depth) - 1) & 1; for (uint32_t w = 0; !mbar_try_wait(&bars[s], par); ++w) if (w > (1u << 24)) return; }
            if (lane == 0) mbar_expect_tx(&bars[s], 32768);
            __syncwarp();
            const int base = ((blockIdx.x * 977 + it * 131) % (nrows_list - 128));
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int g = lane + 32 * u;                 // 64 gather4: 32 row groups x {col 0, col 64}
                const int rg = g & 31, c = (g >> 5) * 64;
                const int *r = rows + base + 4 * rg;
                gather4(smem + s * 32768 + (g >> 5) * 16384 + rg * 512, &tm, c, r[0], r[1], r[2], r[3], &bars[s]);
            }
        }
        for (int it = iters > depth ? iters - depth : 0; it < iters; it++) {
            const int s = it % depth; const uint32_t par = (it / depth) & 1;
            for (uint32_t w = 0; !mbar_try_wait(&bars[s], par); ++w) if (w > (1u << 24)) return;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int R = 4 * 4096 * 4, C = 256;     // 65536 rows x 256 bf16 = 32 MB (L2 resident)
    std::vector<__nv_bfloat16> h((size_t)R * C);
    for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) h[(size_t)r * C + c] = __float2bfloat16((float)((r * 7 + c) % 251));
    __nv_bfloat16 *d; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    void *fp = nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R}; cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64u, 1u}; cuuint32_t estr[2] = {1u, 1u};
    CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode box{64,1}: %d\n", (int)cr);
    if (cr != CUDA_SUCCESS) return 1;

    // correctness
    std::vector<int> rows(128);
    srand(1);
    for (int i = 0; i < 128; i++) rows[i] = rand() % R;
    int *drows; CK(cudaMalloc(&drows, 128 * 4)); CK(cudaMemcpy(drows, rows.data(), 128 * 4, cudaMemcpyHostToDevice));
    uint4 *ddump; CK(cudaMalloc(&ddump, 16384));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480));
    const int col0 = 64;
    probe_kernel<<<1, 128, 20480>>>(tm, drows, col0, ddump);
    CK(cudaDeviceSynchronize());
    std::vector<uint16_t> dump(8192);
    CK(cudaMemcpy(dump.data(), ddump, 16384, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r = 0; r < 128; r++) for (int c = 0; c < 64; c++) {
        const int chunk = (c >> 3) ^ (r & 7);
        const uint16_t got = dump[r * 64 + chunk * 8 + (c & 7)];
        const __nv_bfloat16 e = h[(size_t)rows[r] * C + col0 + c];
        uint16_t ex; memcpy(&ex, &e, 2);
        if (got != ex) { if (bad < 5) printf("mismatch r=%d c=%d got=%04x exp=%04x\n", r, c, got, ex); bad++; }
    }
    printf("gather4 swizzle-128B panel check: %s (%d mismatches)\n", bad ? "FAIL" : "OK", bad);

    // throughput
    const int NL = 1 << 16;
    std::vector<int> list(NL);
    // realistic pattern: runs of ~3 consecutive source pixels (rows) within one pair's 4096-row map, then a jump of ~64
    { int p = 0; for (int i = 0; i < NL; i++) { list[i] = p % R; p += ((i % 3) == 2) ? 62 : 1; } }
    int *dl; CK(cudaMalloc(&dl, NL * 4)); CK(cudaMemcpy(dl, list.data(), NL * 4, cudaMemcpyHostToDevice));
    unsigned long long *dc; CK(cudaMalloc(&dc, 148 * 8));
    const int smem = 6 * 32768 + 1024 + 64;
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int grid : {1, 148}) for (int depth : {1, 2, 3, 4, 6}) {
        const int iters = 512;
        stream_kernel<<<grid, 64, smem>>>(tm, dl, NL, iters, depth, dc);   // warm L2
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        stream_kernel<<<grid, 64, smem>>>(tm, dl, NL, iters, depth, dc);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<unsigned long long> cyc(grid);
        CK(cudaMemcpy(cyc.data(), dc, grid * 8, cudaMemcpyDeviceToHost));
        double avg = 0; for (auto c : cyc) avg += (double)c; avg /= grid;
        const double bytes = (double)iters * 32768;
        printf("grid=%3d depth=%d: %.1f B/clk/SM (clock64), %.1f GB/s chip (events %.3f ms)\n", grid, depth, bytes / avg, bytes * grid / ms * 1e-6, ms);
    }
    return bad ? 1 : 0;
}
