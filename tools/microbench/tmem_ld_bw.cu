// Microbenchmark: tcgen05.ld.32x32b.x32 throughput per SM for 1 / 4 / 8 warps, and MUFU.EX2 throughput.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__global__ void k_tmem(int iters, int unroll_wait, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + (((uint32_t)(warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t r[32];
        tmem_ld32(base + (uint32_t)((i * 32) & 255) + ((warp >> 2) & 1) * 256, r);
        if (unroll_wait) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j += 8) acc ^= r[j];
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    sink[threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}
__global__ void k_mufu(int iters, long long* cycles, float* sink) {
    float x[8];
    for (int j = 0; j < 8; ++j) x[j] = threadIdx.x * 1e-3f + j;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    float s = 0; for (int j = 0; j < 8; ++j) s += x[j];
    sink[threadIdx.x] = s;
}
int main() {
    long long* c; uint32_t* s; float* fs;
    cudaMalloc(&c, 8); cudaMalloc(&s, 4096); cudaMalloc(&fs, 4096);
    const int iters = 4096;
    for (int w : {1, 4, 8}) for (int wait : {1, 0}) {
        k_tmem<<<1, 32 * w>>>(iters, wait, c, s);
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("tcgen05.ld.32x32b.x32: %d warps, wait-each=%d: %.1f clk per ld per warp, %.1f B/clk per SM (err %s)\n", w, wait,
               (double)h / iters, (double)w * iters * 4096.0 / h, cudaGetErrorString(cudaGetLastError()));
    }
    for (int w : {1, 4, 8, 16}) {
        k_mufu<<<1, 32 * w>>>(iters, c, fs);
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("MUFU.EX2: %d warps: %.2f lane-ops per clk per SM\n", w, (double)w * 32 * 8 * iters / h);
    }
    return 0;
}
