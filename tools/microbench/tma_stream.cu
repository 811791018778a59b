// Microbenchmark: how fast can persistent CTAs stream the attention operands (Q, K, V tiles of one head) with TMA?
//   layout 0: the QKV GEMM's row-major [rows][3*D] buffer — every tile row is a 128-byte piece at a 3*D*2-byte stride
//   layout 1: head-major [3][B][H][N][64] — every tile is one contiguous block
// A ring of `stages` slots (one box each); the consumer frees a slot as soon as it is full, so the result is the pure
// load rate for that amount of bytes in flight.   usage: tma_stream <layout> <stages> [B=256] [N=197] [H=12]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
struct P { int layout, stages, B, N, H, nkp, rows1; };
constexpr int SLOT = 208 * 128;
__global__ void __launch_bounds__(64, 1) k_stream(const __grid_constant__ CUtensorMap mq0, const __grid_constant__ CUtensorMap mq1,
                                                   const __grid_constant__ CUtensorMap mkv, P p) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t bar = base + p.stages * SLOT;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(bar + 16 * s, 1); mbar_init(bar + 16 * s + 8, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_items = p.B * p.H, D = p.H * 64;
    const int my = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int nbox = my * 4;
    if (threadIdx.x == 0) {
        for (int j = 0; j < nbox; ++j) {
            const int item = blockIdx.x + (j >> 2) * gridDim.x, b = item / p.H, h = item % p.H, which = j & 3;
            const int s = j % p.stages;
            mbar_wait(bar + 16 * s + 8, ((j / p.stages) & 1) ^ 1);
            const CUtensorMap* m = which == 1 ? &mq0 : which == 2 ? &mq1 : &mkv;
            const int rows = which == 1 ? 128 : which == 2 ? p.rows1 : p.nkp;
            const int sec = which == 0 ? 1 : which == 3 ? 2 : 0, r0 = which == 2 ? 128 : 0;
            mbar_expect(bar + 16 * s, rows * 128);
            if (p.layout == 0) tma_load_2d(base + s * SLOT, m, bar + 16 * s, sec * D + h * 64, b * p.N + r0);
            else tma_load_2d(base + s * SLOT, m, bar + 16 * s, 0, ((sec * p.B + b) * p.H + h) * p.N + r0);
        }
    } else if (threadIdx.x == 32) {
        for (int j = 0; j < nbox; ++j) {
            const int s = j % p.stages;
            mbar_wait(bar + 16 * s, (j / p.stages) & 1);
            mbar_arrive(bar + 16 * s + 8);
        }
    }
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap mk(EncodeFn fn, void* ptr, long rows, long cols, int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}
int main(int argc, char** argv) {
    P p;
    p.layout = argc > 1 ? atoi(argv[1]) : 0; p.stages = argc > 2 ? atoi(argv[2]) : 4;
    p.B = argc > 3 ? atoi(argv[3]) : 256; p.N = argc > 4 ? atoi(argv[4]) : 197; p.H = argc > 5 ? atoi(argv[5]) : 12;
    p.nkp = (p.N + 15) / 16 * 16; p.rows1 = (p.N - 128 + 7) / 8 * 8;
    const long rows = (long)p.B * p.N, D = p.H * 64;
    void* buf; cudaMalloc(&buf, rows * 3 * D * 2 + (1 << 20)); cudaMemset(buf, 0, rows * 3 * D * 2);
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeFn fn = (EncodeFn)fnp;
    const long mrows = p.layout == 0 ? rows : rows * 3 * p.H, mcols = p.layout == 0 ? 3 * D : 64;
    CUtensorMap mq0 = mk(fn, buf, mrows, mcols, 128), mq1 = mk(fn, buf, mrows, mcols, p.rows1), mkv = mk(fn, buf, mrows, mcols, p.nkp);
    const int smem = p.stages * SLOT + 16 * p.stages + 2048;
    cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(e0);
        k_stream<<<sms, 64, smem>>>(mq0, mq1, mkv, p);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
    }
    const double bytes = (double)p.B * p.H * (128 + p.rows1 + 2 * p.nkp) * 128;
    printf("layout %d stages %d (%.0f KB in flight/SM) B=%d N=%d H=%d: %.1f us  %.0f GB/s   (%s)\n", p.layout, p.stages, p.stages * SLOT / 1024.0, p.B, p.N, p.H,
           best * 1e3, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
