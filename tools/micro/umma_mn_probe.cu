// Probe (development aid): shared-memory descriptor of an MN-major B operand for tcgen05.mma kind::f16.
// D[128 x 32] = A[128 x 128] (bf16, K-major, SWIZZLE_64B k-blocks of 32) * V[128 keys x 32] (bf16, rows = K index, 64 B per row:
// N contiguous = "MN-major"), i.e. the P V product of window attention without transposing V. Tries (LBO, SBO, K-advance) variants.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_mn_probe tools/micro/umma_mn_probe.cu && /tmp/umma_mn_probe
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;   // 2 = SW128, 4 = SW64, 6 = SW32, 0 = none
    return d;
}
__device__ __forceinline__ int sw64_off(int r, int c) {   // element (row r, col c < 32) of a [rows x 32] bf16 tile, 64 B rows
    return r * 64 + ((((c >> 3) ^ (r >> 1)) & 3) << 4) + (c & 7) * 2;
}

__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16* A, const __nv_bfloat16* V, float* D, uint32_t lbo, uint32_t sbo, uint32_t kadv,
                                                int swz_v) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar;
    uint8_t* sa = smem;             // 4 k-blocks x [128 x 64 B]
    uint8_t* sv = smem + 32768;     // [128 keys x 64 B]
    const int tid = threadIdx.x;
    for (int i = tid; i < 128 * 128; i += 128) {
        const int r = i / 128, c = i % 128;
        *reinterpret_cast<__nv_bfloat16*>(sa + (c >> 5) * 8192 + sw64_off(r, c & 31)) = A[i];
    }
    for (int i = tid; i < 128 * 32; i += 128) {
        const int k = i / 32, n = i % 32;
        const int off = swz_v ? sw64_off(k, n) : k * 64 + n * 2;
        *reinterpret_cast<__nv_bfloat16*>(sv + off) = V[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tptr;
    if (tid == 0) {
        // idesc: c_format F32 (bit 4), a/b BF16 (bits 7, 10), b_major = MN (bit 16), N >> 3 at 17, M >> 4 at 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        for (int ks = 0; ks < 8; ++ks) {
            const uint64_t da = desc(smem_u32(sa) + (ks >> 1) * 8192 + (ks & 1) * 32, 16, 512, 4);
            const uint64_t db = desc(smem_u32(sv) + ks * kadv, lbo, sbo, swz_v ? 4 : 0);
            const uint32_t acc = ks != 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                         ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    const uint32_t a = tm + ((uint32_t)((tid >> 5) * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(a) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[tid * 32 + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32) : "memory");
}

int main() {
    std::vector<__nv_bfloat16> A(128 * 128), V(128 * 32);
    std::vector<float> Af(128 * 128), Vf(128 * 32), ref(128 * 32, 0.f), got(128 * 32);
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 9) & 0xFFFF) / 65536.0f - 0.5f; };
    for (int i = 0; i < 128 * 128; ++i) { A[i] = __float2bfloat16(rnd()); Af[i] = __bfloat162float(A[i]); }
    for (int i = 0; i < 128 * 32; ++i) { V[i] = __float2bfloat16(rnd()); Vf[i] = __bfloat162float(V[i]); }
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
            float a = 0.f;
            for (int k = 0; k < 128; ++k) a += Af[m * 128 + k] * Vf[k * 32 + n];
            ref[m * 32 + n] = a;
        }
    __nv_bfloat16 *dA, *dV; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dV, V.size() * 2); cudaMalloc(&dD, got.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dV, V.data(), V.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    struct Var { uint32_t lbo, sbo, kadv; int swz; };
    const Var vars[] = {{16, 512, 1024, 1}, {512, 512, 1024, 1}, {512, 16, 1024, 1}, {1024, 512, 1024, 1}, {512, 1024, 1024, 1},
                        {64, 512, 1024, 1}, {512, 64, 1024, 1}, {16, 512, 1024, 0}, {512, 512, 1024, 0}, {64, 1024, 1024, 0}, {1024, 64, 1024, 0},
                        {128, 1024, 1024, 0}, {1024, 128, 1024, 0}};
    for (const Var& v : vars) {
        cudaMemset(dD, 0, got.size() * 4);
        probe<<<1, 128, 44 * 1024>>>(dA, dV, dD, v.lbo, v.sbo, v.kadv, v.swz);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("lbo %u sbo %u kadv %u swz %d: CUDA error %s\n", v.lbo, v.sbo, v.kadv, v.swz, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost);
        double num = 0, den = 0;
        for (size_t i = 0; i < got.size(); ++i) { num += (got[i] - ref[i]) * (double)(got[i] - ref[i]); den += ref[i] * (double)ref[i]; }
        printf("lbo %4u sbo %4u kadv %4u swizzle64 %d: rel err %.3e %s\n", v.lbo, v.sbo, v.kadv, v.swz, sqrt(num / den), sqrt(num / den) < 1e-3 ? "<== MATCH" : "");
    }
    return 0;
}
