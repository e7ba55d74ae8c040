// Micro-benchmark: tcgen05.ld / MUFU.EX2 throughput per SM on sm_100a (design input for the attention kernels).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* out, float* sink) {
    __shared__ uint32_t tb;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tb)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    long long t0 = clock64();
    if (MODE == 0) {           // tcgen05.ld 32x32b.x32 back to back, one wait per 4 loads
        for (int i = 0; i < iters; ++i) {
            uint32_t r[32];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(base + c * 32) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += __uint_as_float(r[0] & 0x3f800000u);
            }
        }
    } else if (MODE == 1) {    // MUFU.EX2: 8 independent chains
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = -0.001f * (threadIdx.x + j);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int c = 0; c < 16; ++c)
#pragma unroll
                for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += x[j];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}

int main() {
    long long* out; float* sink;
    cudaMallocManaged(&out, 1024 * 8); cudaMalloc(&sink, 4);
    const int iters = 2000;
    for (int warps : {4, 8, 16}) {
        k<0><<<148, warps * 32>>>(iters, out, sink);
        cudaDeviceSynchronize();
        double clk = (double)out[0];
        double bytes = (double)warps * iters * 4 * 32 * 32 * 4;
        printf("tcgen05.ld 32x32b.x32  %2d warps: %.0f clk, %.1f B/clk/SM\n", warps, clk, bytes / clk);
        k<1><<<148, warps * 32>>>(iters, out, sink);
        cudaDeviceSynchronize();
        clk = (double)out[0];
        double ops = (double)warps * 32 * iters * 16 * 8;
        printf("ex2.approx             %2d warps: %.0f clk, %.2f ex2/clk/SM\n", warps, clk, ops / clk);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
