// Micro-benchmark (development aid): tensor-memory read / write throughput of one SM, and MUFU.EX2 rate for reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ldtm_rate tools/micro/ldtm_rate.cu && build/ldtm_rate
// Each of W warps (quadrant = warp & 3) issues `iters` tcgen05.ld.32x32b.x32 over rotating columns; bytes = W * 32 * 32 * 4 * iters.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int MODE>   // 0: ld x32 + wait each; 1: 4 ld x32 in flight then wait; 2: st x16; 3: ld x16
__global__ void __launch_bounds__(512, 1) k(long long* out, int iters, int nwarps) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nwarps) {
        for (int i = 0; i < iters; ++i) {
            if (MODE == 0 || MODE == 1) {
                const int n = MODE == 0 ? 1 : 4;
                uint32_t v[4][32];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q < n) {
                        const uint32_t a = base + (((i * 4 + q + warp) * 32) & 511 & ~31);
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                            : "=r"(v[q][0]), "=r"(v[q][1]), "=r"(v[q][2]), "=r"(v[q][3]), "=r"(v[q][4]), "=r"(v[q][5]), "=r"(v[q][6]), "=r"(v[q][7]),
                              "=r"(v[q][8]), "=r"(v[q][9]), "=r"(v[q][10]), "=r"(v[q][11]), "=r"(v[q][12]), "=r"(v[q][13]), "=r"(v[q][14]), "=r"(v[q][15]),
                              "=r"(v[q][16]), "=r"(v[q][17]), "=r"(v[q][18]), "=r"(v[q][19]), "=r"(v[q][20]), "=r"(v[q][21]), "=r"(v[q][22]), "=r"(v[q][23]),
                              "=r"(v[q][24]), "=r"(v[q][25]), "=r"(v[q][26]), "=r"(v[q][27]), "=r"(v[q][28]), "=r"(v[q][29]), "=r"(v[q][30]), "=r"(v[q][31])
                            : "r"(a) : "memory");
                    }
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (q < n) acc += v[q][0] ^ v[q][31];
            } else if (MODE == 2) {
                const uint32_t a = base + (((i + warp) * 16) & 511 & ~15);
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(a), "r"(acc) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = acc; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    const int iters = 2000;
    for (int mode = 0; mode < 3; ++mode)
        for (int nw : {4, 8, 16}) {
            long long h[2];
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, 512>>>(d, iters, nw);
                if (mode == 1) k<1><<<148, 512>>>(d, iters, nw);
                if (mode == 2) k<2><<<148, 512>>>(d, iters, nw);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            const double per = mode == 2 ? 32.0 * 16 * 4 : (mode == 0 ? 1 : 4) * 32.0 * 32 * 4;
            printf("mode %d (%s) warps %2d: %lld cycles, %.1f B/clk/SM, %.1f cycles per warp-op\n", mode,
                   mode == 0 ? "ld.x32 + wait" : mode == 1 ? "4 x ld.x32 + wait" : "st.x16 + wait", nw, h[0], per * nw * iters / (double)h[0],
                   (double)h[0] / iters / (mode == 1 ? 4 : 1));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
