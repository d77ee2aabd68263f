// Microbenchmark: per-SM throughput of the MUFU variants the GELU epilogues can use (development aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mufu_rate tools/micro/mufu_rate.cu && build/mufu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
template <int OP>
__global__ void k(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(a[i]));
            if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
            if (OP == 2) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(*reinterpret_cast<float*>(&a[i])));
            if (OP == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(*reinterpret_cast<float*>(&a[i])));
            if (OP == 4) asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(a[i]));
            if (OP == 5) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(a[i]));
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const char* names[] = {"tanh.approx.f16x2", "ex2.approx.f16x2", "tanh.approx.f32", "ex2.approx.ftz.f32", "fma.rn.f16x2", "tanh.approx.bf16x2"};
    for (int op = 0; op < 6; ++op) {
        for (int threads : {128, 512, 1024}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (op == 0) k<0><<<148, threads>>>(out, cyc, 1); if (op == 1) k<1><<<148, threads>>>(out, cyc, 1);
                if (op == 2) k<2><<<148, threads>>>(out, cyc, 1); if (op == 3) k<3><<<148, threads>>>(out, cyc, 1);
                if (op == 4) k<4><<<148, threads>>>(out, cyc, 1); if (op == 5) k<5><<<148, threads>>>(out, cyc, 1);
                cudaDeviceSynchronize();
            }
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            double insts = (double)ITER * 8 * threads;   // thread-level PTX instructions per SM
            printf("%-20s threads/SM %4d: %.2f thread-instr/clk/SM  (%.2f elements/clk/SM)\n", names[op], threads, insts / c,
                   insts / c * ((op == 0 || op == 1 || op == 4 || op == 5) ? 2 : 1));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
