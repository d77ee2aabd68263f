// Dual tcgen05 GEMM of the training backward: TWO products with the same shape share one tile schedule and meet in the epilogue,
// so neither result is ever written to HBM.
//
//   acc1[M,N] = A1[M,K] W1[N,K]^T      acc2[M,N] = A2[M,K] W2[N,K]^T        (bf16 operands, K-major, fp32 accumulators in TMEM)
//
//   MODE 0 (FFN backward, reference: autograd of Mlp, htsat.py:146-164 through src/training.py:30):
//          out = bf16( acc1 * gelu'(acc2 + bias2) )        A1 = bf16(dL/dy), W1 = fc2.weight^T [4C, C]   -> g W2
//                                                          A2 = LayerNorm(x), W2 = fc1.weight  [4C, C]   -> the recomputed pre-activation
//          replaces the fc1 re-computation GEMM (bf16 hpre to HBM), and the "(g W2) * gelu'(hpre)" GEMM that read it back
//          (stage 0: 228 + 600 us, 1.6 GB of HBM traffic for hpre alone).
//   MODE 1 (lambda gradient, reference: autograd of ResiDual.forward, src/residual.py:37-40):
//          coef = acc1 + bias1 (= x_proj),  gcoef = acc2 (= dL/d x_scaled)
//          dlam[n] += sum_m coef * gcoef   (n < Kvalid),     out = bf16(gcoef * lam[n])
//          replaces two fp32-output GEMMs (coef, gcoef to HBM) and the lambda_grad kernel that read both back.
//
// Structure (one CTA per SM, 18 warps): warp 0 = TMA producer (four 128 x 64 SWIZZLE_128B tiles per k-block stage, 2 stages),
// warp 1 = TMEM allocator + MMA issuer (M128 N128 K16 `tcgen05.mma`, two accumulators per tile, two tile buffers = all 512
// columns), warps 2-17 = epilogue: warp (quadrant q, group g) owns rows 32q..32q+31 and columns 32g..32g+31 of every tile,
// `tcgen05.ld` of both accumulators, packed-fp16 gelu' (ard_common.cuh) or the lambda product, swizzled staging tile, TMA store.
// A CTA keeps ONE column block for its whole life (tiles are strided over row blocks only), so the per-column vectors sit in
// shared memory once and MODE 1 accumulates its column sums in registers across all of the CTA's tiles: one shared-memory
// reduction and 128 global atomics per CTA at the very end.
#include "ard_common.cuh"
#include "ard_internal.h"

namespace ard {

namespace dual {
constexpr int BM = 128, BN = 128, BK = 64;
constexpr int NEPI = 16;
constexpr int THREADS = 64 + NEPI * 32;
constexpr int TILE_BYTES = BM * BK * 2;        // one 128 x 64 bf16 operand tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES;    // A1, W1, A2, W2
constexpr int STAGES = 2;
constexpr int CST = 2048;                      // one 32 x 32 bf16 staging chunk
constexpr int CSTAGE_OFF = STAGES * STAGE_BYTES;
constexpr int VEC_OFF = CSTAGE_OFF + NEPI * 2 * CST;   // [2][128] floats: per-column bias / lambda of this CTA's column block
constexpr int RED_OFF = VEC_OFF + 2 * BN * 4;          // [128] floats: MODE 1 column sums
constexpr int BAR_OFF = RED_OFF + BN * 4;
constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "gemm_dual: shared memory budget");
}  // namespace dual

struct DualParams {
    int M, N, K;
    const float* vec1;   // MODE 0: bias2 [N] (fc1 bias);  MODE 1: bias1 [N] (c0)
    const float* vec2;   // MODE 1: lambda [N] (zero beyond Kvalid)
    float* dlam;         // MODE 1: [Kvalid] accumulated with atomics (may be null)
    int Kvalid;
};

template <int MODE>
__global__ void __launch_bounds__(dual::THREADS, 1)
gemm_dual_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmC, const DualParams p) {
    using namespace dual;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* vec = reinterpret_cast<float*>(smem + VEC_OFF);
    float* red = reinterpret_cast<float*>(smem + RED_OFF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blocks = (p.M + BM - 1) / BM;
    const int n_blocks = (p.N + BN - 1) / BN;
    const int n_blk = (int)blockIdx.x % n_blocks;          // this CTA's column block, fixed
    const int m_first = (int)blockIdx.x / n_blocks;
    const int m_step = (int)gridDim.x / n_blocks;
    const int num_kb = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmC);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], NEPI); }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < BN; i += THREADS) {
        const int col = n_blk * BN + i;
        vec[i] = (p.vec1 != nullptr && col < p.N) ? p.vec1[col] : 0.0f;
        vec[BN + i] = (p.vec2 != nullptr && col < p.N) ? p.vec2[col] : 0.0f;
        red[i] = 0.0f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================================================== TMA producer
        int stage = 0;
        uint32_t phase = 0;
        for (int m_blk = m_first; m_blk < m_blocks; m_blk += m_step) {
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* s0 = smem + stage * STAGE_BYTES;
                if (elect_one_sync()) {
                    mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    tma_load_2d(s0, &tmA1, &full_bar[stage], kb * BK, m_blk * BM);
                    tma_load_2d(s0 + TILE_BYTES, &tmW1, &full_bar[stage], kb * BK, n_blk * BN);
                    tma_load_2d(s0 + 2 * TILE_BYTES, &tmA2, &full_bar[stage], kb * BK, m_blk * BM);
                    tma_load_2d(s0 + 3 * TILE_BYTES, &tmW2, &full_bar[stage], kb * BK, n_blk * BN);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (warp-convergent loop, one elected lane issues)
        const uint32_t idesc = umma_idesc_bf16(BM, BN);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int m_blk = m_first; m_blk < m_blocks; m_blk += m_step, ++it) {
            const int as = it & 1;
            mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d1 = tmem_base + as * 256, d2 = d1 + BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t s0 = smem_u32(smem + stage * STAGE_BYTES);
                const uint64_t da1 = umma_desc_sw128(s0), dw1 = umma_desc_sw128(s0 + TILE_BYTES);
                const uint64_t da2 = umma_desc_sw128(s0 + 2 * TILE_BYTES), dw2 = umma_desc_sw128(s0 + 3 * TILE_BYTES);
                const int ksteps = min(BK, p.K - kb * BK) >> 4;   // 4, or 1..3 in the last k-block (K % 16 == 0)
                if (elect_one_sync()) {
                    if (ksteps == 4) {
                        umma_f16_ss_run<4>(d1, da1, dw1, idesc, kb != 0);
                        umma_f16_ss_run<4>(d2, da2, dw2, idesc, kb != 0);
                    } else {
                        if (ksteps >= 2) {
                            umma_f16_ss_run<2>(d1, da1, dw1, idesc, kb != 0);
                            umma_f16_ss_run<2>(d2, da2, dw2, idesc, kb != 0);
                        }
                        if (ksteps & 1) {
                            const int k = ksteps - 1;
                            umma_bf16_ss(d1, da1 + 2 * k, dw1 + 2 * k, idesc, (kb | k) != 0);
                            umma_bf16_ss(d2, da2 + 2 * k, dw2 + 2 * k, idesc, (kb | k) != 0);
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kb == num_kb - 1) umma_commit(&tfull_bar[as]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================================================== epilogue warps
        const int ew = warp - 2;
        const int quad = warp & 3;        // TMEM lane quadrant this warp may access
        const int c = ew >> 2;            // its 32-column chunk of every tile
        uint8_t* cst = smem + CSTAGE_OFF + ew * 2 * CST;
        const float* v1s = vec + c * 32;          // MODE 0: bias2; MODE 1: bias1 (c0)
        const float* v2s = vec + BN + c * 32;     // MODE 1: lambda
        const int col0 = n_blk * BN + c * 32;
        float psum[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) psum[j] = 0.0f;
        int it = 0;
        if (col0 < p.N) {
            for (int m_blk = m_first; m_blk < m_blocks; m_blk += m_step, ++it) {
                const int as = it & 1;
                mbar_wait(&tfull_bar[as], (it >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + as * 256 + c * 32 + ((uint32_t)(quad * 32) << 16);
                uint8_t* sbuf = cst + (it & 1) * CST;
                if (lane == 0) tma_store_wait_read<1>();   // this buffer was handed to TMA two tiles ago
                __syncwarp();
                uint8_t* rowp = sbuf + lane * 64;          // row = 64 B, CU_TENSOR_MAP_SWIZZLE_64B: 16-byte unit index ^= (row >> 1) & 3
                const int sw = (lane >> 1) & 3;
                // eight columns j0..j0+7 of this thread's row: av = acc1, bv = acc2 -> one 16-byte unit of the staging row
                auto unit = [&](int j0, const uint32_t* av, const uint32_t* bv) {
                    uint32_t o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int j = j0 + 2 * k;
                        const float2 vb = *reinterpret_cast<const float2*>(v1s + j);
                        if constexpr (MODE == 0) {
                            const float2 gd =
                                gelu_erf_grad_h2(__floats2half2_rn(__uint_as_float(bv[2 * k]) + vb.x, __uint_as_float(bv[2 * k + 1]) + vb.y));
                            o[k] = pack_bf16x2(__uint_as_float(av[2 * k]) * gd.x, __uint_as_float(av[2 * k + 1]) * gd.y);
                        } else {
                            const float2 lm = *reinterpret_cast<const float2*>(v2s + j);
                            const float g0 = __uint_as_float(bv[2 * k]), g1 = __uint_as_float(bv[2 * k + 1]);
                            psum[j] = fmaf(__uint_as_float(av[2 * k]) + vb.x, g0, psum[j]);
                            psum[j + 1] = fmaf(__uint_as_float(av[2 * k + 1]) + vb.y, g1, psum[j + 1]);
                            o[k] = pack_bf16x2(g0 * lm.x, g1 * lm.y);
                        }
                    }
                    *reinterpret_cast<uint4*>(rowp + (((j0 >> 3) ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                };
                if constexpr (MODE == 0) {
                    uint32_t a[32], b[32];
                    tmem_ld_32x32b_x32(taddr, a);
                    tmem_ld_32x32b_x32(taddr + BN, b);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[as]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) unit(q * 8, a + q * 8, b + q * 8);
                } else {   // the column sums live in 32 registers: the accumulators are read 16 columns at a time
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t a[16], b[16];
                        tmem_ld_32x32b_x16(taddr + hf * 16, a);
                        tmem_ld_32x32b_x16(taddr + BN + hf * 16, b);
                        tmem_ld_wait();
                        if (hf == 1) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tempty_bar[as]);
                        }
                        unit(hf * 16, a, b);
                        unit(hf * 16 + 8, a + 8, b + 8);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmC, sbuf, col0, m_blk * BM + quad * 32);
                    tma_store_commit();
                }
            }
            if (lane == 0) tma_store_wait_all<0>();
        } else {
            // a column chunk entirely beyond N (N % 128 != 0): only keep the accumulator hand-off going
            for (int m_blk = m_first; m_blk < m_blocks; m_blk += m_step, ++it) {
                const int as = it & 1;
                mbar_wait(&tfull_bar[as], (it >> 1) & 1);
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[as]);
            }
        }
        if constexpr (MODE == 1) {
            // column sums: rows of this warp (lanes) first, then the four quadrant warps of the chunk through shared memory
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float s = psum[j];
                s += __shfl_xor_sync(0xffffffffu, s, 16);
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                if (lane == j) atomicAdd(&red[c * 32 + j], s);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (MODE == 1) {
        if (p.dlam != nullptr && (int)threadIdx.x < BN) {
            const int col = n_blk * BN + (int)threadIdx.x;
            if (col < p.Kvalid && m_first < m_blocks) atomicAdd(p.dlam + col, red[threadIdx.x]);
        }
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int MODE>
static int launch_dual(const DualArgs& a, int num_sms, cudaStream_t stream) {
    using namespace dual;
    if (a.M <= 0 || a.N <= 0 || a.K <= 0 || (a.K % 16) != 0 || (a.lda1 % 8) || (a.lda2 % 8) || (a.ldw1 % 8) || (a.ldw2 % 8) || (a.ldo % 8))
        return set_error(ARD_ERR_SHAPE, "gemm_dual: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
    const int n_blocks = (a.N + BN - 1) / BN;
    if (n_blocks > num_sms) return set_error(ARD_ERR_SHAPE, "gemm_dual: N=%d needs more column blocks than SMs", a.N);
    CUtensorMap t1, w1, t2, w2, tc;
    ARD_TRY(make_tmap_2d(&t1, a.A1, 2, a.K, a.M, (uint64_t)a.lda1 * 2, BK, BM, 128));
    ARD_TRY(make_tmap_2d(&w1, a.W1, 2, a.K, a.N, (uint64_t)a.ldw1 * 2, BK, BN, 128));
    ARD_TRY(make_tmap_2d(&t2, a.A2, 2, a.K, a.M, (uint64_t)a.lda2 * 2, BK, BM, 128));
    ARD_TRY(make_tmap_2d(&w2, a.W2, 2, a.K, a.N, (uint64_t)a.ldw2 * 2, BK, BN, 128));
    ARD_TRY(make_tmap_2d(&tc, a.out, 2, a.N, a.M, (uint64_t)a.ldo * 2, 32, 32, 64));
    auto kern = gemm_dual_kernel<MODE>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return set_error(ARD_ERR_CUDA, "cudaFuncSetAttribute(gemm_dual smem=%d): %s", SMEM_BYTES, cudaGetErrorString(e));
        attr_set = true;
    }
    DualParams p;
    p.M = a.M; p.N = a.N; p.K = a.K; p.vec1 = a.vec1; p.vec2 = a.vec2; p.dlam = a.dlam; p.Kvalid = a.Kvalid;
    const int m_blocks = (a.M + BM - 1) / BM;
    int per_col = num_sms / n_blocks;
    if (per_col > m_blocks) per_col = m_blocks;
    const int grid = per_col * n_blocks;
    ProfScope ps(PROF_GEMM, stream, 4.0 * a.M * a.N * a.K, 4.0 * a.M * a.K + 4.0 * a.N * a.K + 2.0 * a.M * a.N);
    kern<<<grid, THREADS, SMEM_BYTES, stream>>>(t1, w1, t2, w2, tc, p);
    return check_cuda(cudaGetLastError(), "gemm_dual launch");
}

int gemm_dual_gelu_bwd(const DualArgs& a, int num_sms, cudaStream_t stream) { return launch_dual<0>(a, num_sms, stream); }
int gemm_dual_lambda(const DualArgs& a, int num_sms, cudaStream_t stream) { return launch_dual<1>(a, num_sms, stream); }

}  // namespace ard
