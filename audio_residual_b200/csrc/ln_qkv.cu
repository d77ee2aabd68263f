// Fused norm1 + qkv projection of the 96-channel stage:  qkv[M, 288] (bf16) = LayerNorm(x[M, 96]) Wqkv^T + b
//
// Reference: SwinTransformerBlock.forward htsat.py:449 (`x = self.norm1(x)`) followed by WindowAttention.forward htsat.py:329
// (`self.qkv(x)`); the q rows of Wqkv / b carry the head_dim^-0.5 scale (htsat.py:295, folded at weight upload).
//
// Unfused, the LayerNorm kernel writes a bf16 copy of the stage's residual stream that the GEMM reads straight back: 4 B per
// token-channel of HBM traffic and a launch for nothing (146 + 189 us at B = 256, both HBM-bound). Here the 55 KB weight matrix
// stays in shared memory, sixteen warps LayerNorm 128-token tiles straight into the swizzled K-major operand layout (two
// buffers, so tile t+1 is normalised while tile t is multiplied and stored), one warp issues three N = 96 tcgen05 MMA groups per
// tile (q, k and v thirds of the output, each with its own full / free barrier so the next tile's MMAs start as soon as a third
// has been drained) and twelve epilogue warps (TMEM lane quadrant x third) add the bias, pack bf16 and TMA-store.
// HBM sees x once (4 B) and qkv once (6 B).
#include "ard_common.cuh"
#include "ard_internal.h"

namespace ard {

constexpr int LQ_C = 96;
constexpr int LQ_N = 3 * LQ_C;          // 288
constexpr int LQ_BM = 128;
constexpr int LQ_EPI_WARPS = 12;        // quadrant (w & 3) x third (w >> 2)
constexpr int LQ_LN_WARPS = 16;
constexpr int LQ_W_MMA = LQ_EPI_WARPS, LQ_W_LN = LQ_W_MMA + 1;
constexpr int LQ_THREADS = (LQ_W_LN + LQ_LN_WARPS) * 32;   // 928

constexpr int LQ_W_KB = LQ_N * 64;              // bytes per 32-wide k-block of W (288 rows x 64 B)
constexpr int LQ_W_OFF = 0;
constexpr int LQ_W_BYTES = 3 * LQ_W_KB;         // 55296
constexpr int LQ_A_KB = LQ_BM * 64;             // 8192
constexpr int LQ_A_OFF = LQ_W_OFF + LQ_W_BYTES; // 55296 = 54 * 1024
constexpr int LQ_A_BYTES = 3 * LQ_A_KB;         // 24576 per buffer, two buffers
constexpr int LQ_CST_OFF = LQ_A_OFF + 2 * LQ_A_BYTES;          // 12 warps x 2 buffers x (32 rows x 64 B)
constexpr int LQ_VEC_OFF = LQ_CST_OFF + LQ_EPI_WARPS * 4096;   // bias[288] gamma[96] beta[96]
constexpr int LQ_BAR_OFF = LQ_VEC_OFF + (LQ_N + 2 * LQ_C) * 4;
constexpr int LQ_SMEM_BYTES = LQ_BAR_OFF + 128 + 1024;

struct LnQkvParams {
    const float* x;        // [M, 96] fp32
    const float* gamma;    // norm1
    const float* beta;
    const float* bias;     // [288]
    int M;
};

__global__ void __launch_bounds__(LQ_THREADS, 1)
ln_qkv_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut, const LnQkvParams p) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, still a shared-space pointer
    float* biass = reinterpret_cast<float*>(smem + LQ_VEC_OFF);
    float* gs = biass + LQ_N;
    float* bs = gs + LQ_C;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LQ_BAR_OFF);
    uint64_t* w_full = bars + 0;
    uint64_t* a_full = bars + 1;     // [2]
    uint64_t* a_free = bars + 3;     // [2]
    uint64_t* acc_full = bars + 5;   // [3] one per third of the accumulator
    uint64_t* acc_free = bars + 8;   // [3]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = (p.M + LQ_BM - 1) / LQ_BM;
    const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    // (bias / LayerNorm affine are weights, never written by a predecessor kernel: safe to read before pdl_wait)
    for (int i = threadIdx.x; i < LQ_N; i += LQ_THREADS) biass[i] = p.bias[i];
    for (int i = threadIdx.x; i < LQ_C; i += LQ_THREADS) {
        gs[i] = p.gamma[i];
        bs[i] = p.beta[i];
    }
    if (warp == LQ_W_MMA && lane == 0) {
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmOut);
        mbar_init(w_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], LQ_LN_WARPS);
            mbar_init(&a_free[i], 1);
        }
        for (int i = 0; i < 3; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_free[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == LQ_W_MMA) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();

    if (warp < LQ_EPI_WARPS) {
        // ============================================================ epilogue: TMEM third -> + bias -> bf16 -> TMA store
        const int quad = warp & 3, third = warp >> 2;
        uint8_t* cst = smem + LQ_CST_OFF + warp * 4096;
        int cbuf = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int tile = blockIdx.x + t * gridDim.x;
            mbar_wait_parked(&acc_full[third], (uint32_t)(t & 1));
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 3; ++c) {                    // 32-column chunks of this third
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + third * 96 + c * 32 + ((uint32_t)(quad * 32) << 16), v);
                tmem_ld_wait();
                if (c == 2) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_free[third]);
                }
                const float* bb = biass + third * 96 + c * 32;
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bb + i);
                    pk[i / 2] = pack_bf16x2(__uint_as_float(v[i]) + b4.x, __uint_as_float(v[i + 1]) + b4.y);
                    pk[i / 2 + 1] = pack_bf16x2(__uint_as_float(v[i + 2]) + b4.z, __uint_as_float(v[i + 3]) + b4.w);
                }
                if (lane == 0) tma_store_wait_read<1>();     // the store issued two chunks ago has read this staging buffer
                __syncwarp();
                uint8_t* sbuf = cst + cbuf * 2048;
                uint8_t* rowp = sbuf + lane * 64;            // row = 64 B (32 bf16), SWIZZLE_64B: 16-byte unit index ^= (row >> 1) & 3
                const int sw = (lane >> 1) & 3;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(rowp + ((q ^ sw) << 4)) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmOut, sbuf, third * 96 + c * 32, tile * LQ_BM + quad * 32);
                    tma_store_commit();
                }
                cbuf ^= 1;
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    } else if (warp == LQ_W_MMA) {
        // ============================================================ weight load + MMA issue (warp-convergent, elected lane)
        if (elect_one_sync()) {
            mbar_expect_tx(w_full, LQ_W_BYTES);
            for (int kb = 0; kb < 3; ++kb)
                for (int hf = 0; hf < 2; ++hf)
                    tma_load_2d(smem + LQ_W_OFF + kb * LQ_W_KB + hf * 144 * 64, &tmW, w_full, kb * 32, hf * 144);
        }
        __syncwarp();
        mbar_wait(w_full, 0);
        constexpr uint32_t idesc = umma_idesc_bf16(LQ_BM, 96);
        const uint64_t dW = umma_desc_sw64(smem_u32(smem + LQ_W_OFF));
        for (int t = 0; t < my_tiles; ++t) {
            const int ab = t & 1;
            mbar_wait_parked(&a_full[ab], (uint32_t)((t >> 1) & 1));
            const uint64_t dA = umma_desc_sw64(smem_u32(smem + LQ_A_OFF + ab * LQ_A_BYTES));
#pragma unroll 1
            for (int third = 0; third < 3; ++third) {
                mbar_wait_parked(&acc_free[third], (uint32_t)((t & 1) ^ 1));
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t d = tmem_base + third * 96;
#pragma unroll
                    for (int kb = 0; kb < 3; ++kb)           // descriptor start-address field is in 16-byte units
                        umma_f16_ss_run<2>(d, dA + (uint64_t)(kb * (LQ_A_KB >> 4)), dW + (uint64_t)(kb * (LQ_W_KB >> 4) + third * ((96 * 64) >> 4)), idesc,
                                           kb != 0);
                    umma_commit(&acc_full[third]);
                    if (third == 2) umma_commit(&a_free[ab]);
                }
                __syncwarp();
            }
        }
    } else {
        // ============================================================ LayerNorm workers: 8 rows per warp, 8 lanes per row
        const int ew = warp - LQ_W_LN;
        const int l8 = lane & 7, rsub = lane >> 3;     // 4 rows per warp instruction
        int a_off[3];                                  // byte offset of this lane's three 8-byte stores for row-group 0
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int c0 = l8 * 12 + q * 4, kb = c0 >> 5, cc = c0 & 31;
            a_off[q] = kb * LQ_A_KB + (ew * 8 + rsub) * 64 + (((cc >> 3) ^ (rsub >> 1)) << 4) + (cc & 7) * 2;
        }
        for (int t = 0; t < my_tiles; ++t) {
            const int tile = blockIdx.x + t * gridDim.x;
            const int ab = t & 1;
            uint8_t* a1 = smem + LQ_A_OFF + ab * LQ_A_BYTES;
            const long long row0 = (long long)tile * LQ_BM + ew * 8;
            float4 v[2][3];
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
                const long long r = row0 + gi * 4 + rsub;
                if (r < p.M) {
                    const float4* xr = reinterpret_cast<const float4*>(p.x + r * LQ_C + l8 * 12);
                    v[gi][0] = __ldg(xr); v[gi][1] = __ldg(xr + 1); v[gi][2] = __ldg(xr + 2);
                } else {
                    v[gi][0] = v[gi][1] = v[gi][2] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (lane < 24) {                               // pull the rows of this warp's tile after next into L2
                const long long nr0 = row0 + 2LL * gridDim.x * LQ_BM;
                if ((nr0 * LQ_C) * 4 + (lane + 1) * 128 <= (long long)p.M * LQ_C * 4)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.x + nr0 * LQ_C) + lane * 128));
            }
            mbar_wait_parked(&a_free[ab], (uint32_t)(((t >> 1) & 1) ^ 1));   // the MMAs that read this buffer two tiles ago have retired
            float sm[2], qv[2];
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
                sm[gi] = 0.f;
#pragma unroll
                for (int q = 0; q < 3; ++q) sm[gi] += (v[gi][q].x + v[gi][q].y) + (v[gi][q].z + v[gi][q].w);
            }
#pragma unroll
            for (int sh = 4; sh > 0; sh >>= 1)
#pragma unroll
                for (int gi = 0; gi < 2; ++gi) sm[gi] += __shfl_xor_sync(0xffffffffu, sm[gi], sh);
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
                const float mean = sm[gi] * (1.0f / LQ_C);
                qv[gi] = 0.f;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    v[gi][q].x -= mean; v[gi][q].y -= mean; v[gi][q].z -= mean; v[gi][q].w -= mean;
                    qv[gi] += (v[gi][q].x * v[gi][q].x + v[gi][q].y * v[gi][q].y) + (v[gi][q].z * v[gi][q].z + v[gi][q].w * v[gi][q].w);
                }
            }
#pragma unroll
            for (int sh = 4; sh > 0; sh >>= 1)
#pragma unroll
                for (int gi = 0; gi < 2; ++gi) qv[gi] += __shfl_xor_sync(0xffffffffu, qv[gi], sh);
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
                const float rstd = rsqrtf(qv[gi] * (1.0f / LQ_C) + 1e-5f);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int c0 = l8 * 12 + q * 4;
                    const float4 gm = *reinterpret_cast<const float4*>(gs + c0);
                    const float4 bt = *reinterpret_cast<const float4*>(bs + c0);
                    uint2 pk;
                    pk.x = pack_bf16x2(fmaf(v[gi][q].x * rstd, gm.x, bt.x), fmaf(v[gi][q].y * rstd, gm.y, bt.y));
                    pk.y = pack_bf16x2(fmaf(v[gi][q].z * rstd, gm.z, bt.z), fmaf(v[gi][q].w * rstd, gm.w, bt.w));
                    // row rr = ew*8 + gi*4 + rsub; SWIZZLE_64B unit index ^= (rr >> 1) & 3 = ((gi & 1) << 1) | (rsub >> 1)
                    *reinterpret_cast<uint2*>(a1 + ((a_off[q] ^ ((gi & 1) << 5)) + gi * 256)) = pk;
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[ab]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == LQ_W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// qkv[M, 288] bf16 = LayerNorm(x; gamma, beta) w^T + bias, C = 96 (w [288, 96] bf16 row-major as nn.Linear stores it).
int ln_qkv_96(const float* x, const float* gamma, const float* beta, const __nv_bfloat16* w, const float* bias, __nv_bfloat16* qkv, long long M,
              int num_sms, cudaStream_t stream) {
    if (M <= 0) return 0;
    if (M > 0x7fffffffLL) return set_error(ARD_ERR_SHAPE, "ln_qkv: too many rows");
    CUtensorMap tw, to;
    ARD_TRY(make_tmap_2d(&tw, w, 2, LQ_C, LQ_N, (uint64_t)LQ_C * 2, 32, 144, 64));
    ARD_TRY(make_tmap_2d(&to, qkv, 2, LQ_N, (uint64_t)M, (uint64_t)LQ_N * 2, 32, 32, 64));
    static bool attr_set = false;
    if (!attr_set) {
        ARD_CUDA(cudaFuncSetAttribute(ln_qkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LQ_SMEM_BYTES));
        attr_set = true;
    }
    LnQkvParams p;
    p.x = x; p.gamma = gamma; p.beta = beta; p.bias = bias; p.M = (int)M;
    const int tiles = (int)((M + LQ_BM - 1) / LQ_BM);
    const int grid = tiles < num_sms ? tiles : num_sms;
    ProfScope ps(PROF_GEMM, stream, 2.0 * M * LQ_N * LQ_C, (double)M * LQ_C * 4.0 + (double)M * LQ_N * 2.0 + 2.0 * LQ_N * LQ_C);
    ARD_CUDA(enqueue_pdl(ln_qkv_kernel, dim3(grid), dim3(LQ_THREADS), LQ_SMEM_BYTES, stream, tw, to, p));
    return check_cuda(cudaGetLastError(), "ln_qkv launch");
}

}  // namespace ard
