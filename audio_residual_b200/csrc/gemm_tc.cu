// Persistent warp-specialised tcgen05 GEMM for sm_100a:   out[M,N] = epilogue( A[M,K] (bf16) * W[N,K]^T (bf16) )
//
// This is the contraction engine behind every Linear on the HTSAT path (reference call sites:
// WindowAttention.qkv / .proj htsat.py:330,354-355; Mlp.fc1 / .fc2 htsat.py:159-163; PatchMerging.reduction htsat.py:524;
// ResiDual projections src/residual.py:38-40; audio_projection model.py:539-543).
//
//   warp 0      : TMA producer  (A tile 128x64 and W tile BNx64 per stage, SWIZZLE_128B, mbarrier complete_tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=BN, K=16 per instruction, fp32 accum in TMEM,
//                 two accumulator buffers so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 2..9  : epilogue. Warp w owns TMEM lanes 32*(w%4)..+31 (one output row per thread), pulls 32-column chunks with
//                 tcgen05.ld, applies bias / exact-erf GELU / ReLU / up to two fp32 residual adds in registers, writes the chunk
//                 into a swizzled shared-memory staging tile and hands it to a TMA store (no global store instructions).
//
// The contraction is tensor-core work; for the small-K layers of stages 0/1 (K = 96/192) the kernel is bound by the
// epilogue's CUDA-core instructions and by HBM, so the epilogue is kept to a few instructions per element.
#include <stdlib.h>

#include "ard_common.cuh"
#include "ard_internal.h"

namespace ard {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;

template <int BN, bool OUT_BF16, bool PAIR = false, bool MUL = false, int NEPI_ = 8>
struct GemmCfg {
    // Epilogue warps: 8, or 16 for the 16-bit-output kernels with a transcendental epilogue (GELU, gelu'). ncu on the stage-2
    // fc1+GELU GEMM with 8 warps: issue slots 52 % busy (two dependent-chain warps per scheduler) while the tensor pipe idled
    // at 37 % - that work needs more warps in flight (measured 105 -> 94 us, stage 1: 189 -> 143 us). The plain 16-bit
    // kernels (qkv, dgrads) are faster with 8 warps and the fourth operand stage the saved staging memory buys.
    static constexpr int NEPI = NEPI_;
    static constexpr int THREADS = 64 + NEPI * 32;
    static constexpr int NGRP = NEPI / 4;                       // warps per TMEM lane quadrant = column-chunk interleave groups
    static constexpr int NCHUNK = BN / 32;
    static constexpr int ACTIVE_GRP = NGRP < NCHUNK ? NGRP : NCHUNK;
    // Staging buffers per epilogue warp: 2 when results are only stored; 3 when a tile is TMA-prefetched INTO the staging
    // buffer one chunk ahead (the fp32 shortcut residual, or the bf16 gelu' multiplicand of the FFN backward, MUL): one
    // buffer is landing, one is being processed, one is still being stored.
    static constexpr int NBUF = (OUT_BF16 && !MUL) ? 2 : 3;
    static constexpr int CST = OUT_BF16 ? 2048 : 4096;          // one 32x32 chunk: 2 KB 16-bit / 4 KB fp32
    // PAIR (cta_group::2): each CTA stages its own 128 rows of A and only HALF of the W tile, so a stage is smaller and the
    // ring is deeper; the L2 -> SM operand traffic per output element drops by a third (256 x BN tile per pair).
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * GEMM_BK * 2;
    // TMEM: 512 fp32 columns = 2 accumulators of up to 256 columns, or 4 of up to 128 (more tiles in flight between the MMA
    // issuer and the epilogue warps for the narrow-N kernels)
    static constexpr int ACC_COLS = BN <= 128 ? 128 : 256;
    static constexpr int NACC = 512 / ACC_COLS;
    static constexpr int STAGES_FIT = (227 * 1024 - 1536 - NEPI * (NBUF * CST + 128)) / (A_BYTES + B_BYTES);
    static constexpr int STAGES = STAGES_FIT > 6 ? 6 : STAGES_FIT;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int CSTAGE_OFF = STAGES * STAGE_BYTES;
    static constexpr int BIAS_OFF = CSTAGE_OFF + NEPI * NBUF * CST;
    static constexpr int BAR_OFF = BIAS_OFF + NEPI * 32 * 4;
    static constexpr int SMEM_BYTES = BAR_OFF + 512 + 1024;  // barriers + alignment slack
};

struct GemmKernelParams {
    int M, N, K;
    const float* bias;   // [N] or null
    int act;             // 0 none, 1 gelu(erf), 2 relu
    const float* resid1; // fp32 [M, ldr1] or null
    long long ldr1;
    const float* resid2;
    long long ldr2;
    float* aux;          // optional fp32 copy of (acc+bias) BEFORE residual adds, row m -> aux[(m / aux_T) * aux_bstride + m % aux_T]
    long long ld_aux;
    int aux_T;
    long long aux_bstride;
    int ab_f16;          // A and W hold fp16 (else bf16)
    int out_f16;         // 16-bit output holds fp16 (GELU evaluated in packed fp16), else bf16
    int mul_gelu_bwd;    // 16-bit output only: out = acc * gelu'(h), h = the bf16 [M,N] tensor behind tmR (FFN backward: dh = (g W2) * gelu'(hpre))
    int splitk;          // >= 1. Split s works on k-blocks [s * kb_per_split, ...) of every tile and stores its partial tile `split_rows`
    int kb_per_split;    // rows further down the output (out is [splitk * split_rows, N] fp32, no bias / residuals): long-K, few-tile
    int split_rows;      // problems (X^T X over a million rows with a 96 x 96 result) would otherwise run on a single CTA
};

template <int BN, bool OUT_BF16, bool PAIR, bool MUL, int NEPI>
__global__ void __launch_bounds__((GemmCfg<BN, OUT_BF16, PAIR, MUL, NEPI>::THREADS), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const GemmKernelParams p) {
    using Cfg = GemmCfg<BN, OUT_BF16, PAIR, MUL, NEPI>;
    constexpr int GEMM_NEPI = Cfg::NEPI;
    constexpr int TILE_M = PAIR ? 2 * GEMM_BM : GEMM_BM;      // rows of one scheduled tile (per CTA pair in PAIR mode)
    const uint32_t pair_rank = PAIR ? cluster_ctarank() : 0;  // 0 = leader (issues the MMAs)
    const int sched_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int sched_n = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, still a shared-space pointer
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    uint64_t* empty_bar = full_bar + Cfg::STAGES;
    uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
    constexpr int NACC = Cfg::NACC;
    uint64_t* tempty_bar = tfull_bar + NACC;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + NACC);
    uint64_t* resid_bar = tempty_bar + NACC + 1;   // [NEPI][3]: prefetched tile landed in staging buffer b

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_blocks = (p.M + TILE_M - 1) / TILE_M;
    const int n_blocks = (p.N + BN - 1) / BN;
    const int mn_tiles = m_blocks * n_blocks;
    const int num_tiles = mn_tiles * p.splitk;
    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        for (int i = 0; i < Cfg::STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < NACC; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], (PAIR ? 2 : 1) * 4 * Cfg::ACTIVE_GRP);
        }
        for (int i = 0; i < GEMM_NEPI * 3; ++i) mbar_init(&resid_bar[i], 1);
        tma_prefetch_desc(&tmR);
        fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (PAIR) {
            tmem_alloc_pair(tmem_ptr_smem, 512);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_ptr_smem, 512);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();   // operands / residuals come from the predecessor kernel; everything above overlapped its tail

    if (warp == 0) {
        // ===================================================== TMA producer (warp-convergent loop, one elected lane issues)
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = sched_id; tile < num_tiles; tile += sched_n) {
                const int ks = tile / mn_tiles, t2 = tile - ks * mn_tiles;
                const int m_blk = t2 / n_blocks, n_blk = t2 % n_blocks;
                const int kb_begin = ks * p.kb_per_split, kb_end = min(num_kb, kb_begin + p.kb_per_split);
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    if (elect_one_sync()) {
                        if constexpr (PAIR) {
                            // both CTAs load their halves; all bytes are counted on the leader's barrier, armed by the leader
                            if (pair_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                            tma_load_2d_pair(sa, &tmA, &full_bar[stage], kb * GEMM_BK, m_blk * TILE_M + (int)pair_rank * GEMM_BM);
                            tma_load_2d_pair(sb, &tmB, &full_bar[stage], kb * GEMM_BK, n_blk * BN + (int)pair_rank * (BN / 2));
                        } else {
                            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                            tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
                            tma_load_2d(sb, &tmB, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
                        }
                    }
                    __syncwarp();
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        // The whole warp runs the loop and one elected lane issues: in warp-convergent code the descriptors, TMEM address and
        // loop state stay on the uniform datapath and a k-block's UTCHMMAs are emitted back to back. Issued from a lane-0 branch
        // every MMA cost ~12 SASS instructions (R2UR per operand, ELECT / BRA.U.ANY wrapper), ~85 cycles on a busy scheduler:
        // more than a BN <= 128 MMA takes to execute.
        if (pair_rank == 0) {
            const uint32_t idesc = p.ab_f16 ? umma_idesc_f16(TILE_M, BN) : umma_idesc_bf16(TILE_M, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = sched_id; tile < num_tiles; tile += sched_n, ++it) {
                const int as = it % NACC;
                const uint32_t aphase = (it / NACC) & 1;
                if constexpr (PAIR) mbar_wait_cluster(&tempty_bar[as], aphase ^ 1);
                else mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * Cfg::ACC_COLS;
                const int kb_begin = (tile / mn_tiles) * p.kb_per_split, kb_end = min(num_kb, kb_begin + p.kb_per_split);
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t da = umma_desc_sw128(sa);
                    const uint64_t db = umma_desc_sw128(sa + Cfg::A_BYTES);
                    const int ksteps = min(GEMM_BK, p.K - kb * GEMM_BK) >> 4;   // 4, or 1..3 in the last k-block (K % 16 == 0)
                    if (elect_one_sync()) {
                        if (ksteps == 4) {
                            umma_f16_ss_run<4, PAIR>(d_tmem, da, db, idesc, kb != kb_begin);
                        } else {
                            if (ksteps >= 2) umma_f16_ss_run<2, PAIR>(d_tmem, da, db, idesc, kb != kb_begin);
                            if (ksteps & 1) {
                                const int k = ksteps - 1;
                                if constexpr (PAIR) umma_bf16_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, ((kb - kb_begin) | k) != 0);
                                else umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, ((kb - kb_begin) | k) != 0);
                            }
                        }
                        // frees the smem slot (in both CTAs of a pair) when these MMAs retire
                        if constexpr (PAIR) umma_commit_pair(&empty_bar[stage]);
                        else umma_commit(&empty_bar[stage]);
                        // accumulator ready (each CTA's epilogue reads its own 128 rows out of its own TMEM)
                        if (kb == kb_end - 1) {
                            if constexpr (PAIR) umma_commit_pair(&tfull_bar[as]);
                            else umma_commit(&tfull_bar[as]);
                        }
                    }
                    __syncwarp();
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================================================== epilogue warps
        const int ew = warp - 2;
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
        const int grp = ew >> 2;                   // column-chunk interleave group
        constexpr int NGRP = Cfg::NGRP;
        constexpr int NCHUNK = Cfg::NCHUNK;
        constexpr int NBUF = Cfg::NBUF;
        uint8_t* cst = smem + Cfg::CSTAGE_OFF + ew * NBUF * Cfg::CST;
        float* bias_w = reinterpret_cast<float*>(smem + Cfg::BIAS_OFF) + ew * 32;
        uint64_t* rbar = resid_bar + ew * 3;
        // A tile the epilogue combines with the accumulator is fetched by TMA into the staging buffer one chunk ahead:
        // coalesced, asynchronous, no registers. (Row-per-thread LDG.128 touches 32 different lines per instruction and made
        // the residual GEMMs LSU-bound.) fp32 output: the shortcut residual, added in place. 16-bit output (MUL): the bf16
        // pre-activation whose gelu' multiplies the result (FFN backward). The same buffer is then handed to the TMA store.
        const bool tma_resid = MUL || (!OUT_BF16 && p.resid1 != nullptr);
        auto issue_resid = [&](int t, int cc, int b) {   // lane 0 only
            const int t2 = t % mn_tiles;
            const int mb = t2 / n_blocks, nb = t2 % n_blocks;
            mbar_expect_tx(&rbar[b], Cfg::CST);
            tma_load_2d(cst + b * Cfg::CST, &tmR, &rbar[b], nb * BN + cc * 32, mb * TILE_M + (int)pair_rank * GEMM_BM + quad * 32);
        };
        int tile = grp < NCHUNK ? sched_id : num_tiles;   // BN = 96 has 3 chunks: the fourth group of a 16-warp epilogue idles
        int c = grp, it = 0, i = 0;
        if (tma_resid && tile < num_tiles && lane == 0) issue_resid(tile, c, 0);
        while (tile < num_tiles) {
            int ntile = tile, nc = c + NGRP;
            if (nc >= NCHUNK) { nc = grp; ntile = tile + sched_n; }
            const int b = i % NBUF;
            if (tma_resid && lane == 0) {
                tma_store_wait_read<1>();          // buffer (i+1)%3 was last stored two chunks ago
                if (ntile < num_tiles) issue_resid(ntile, nc, (i + 1) % NBUF);
            }
            const int ks = tile / mn_tiles, t2 = tile - ks * mn_tiles;
            const int m_blk = t2 / n_blocks, n_blk = t2 % n_blocks;
            const int as = it % NACC;
            if (c == grp) {                        // first chunk of this tile for this warp
                mbar_wait(&tfull_bar[as], (it / NACC) & 1);
                tc_fence_after();
            }
            const int row0 = m_blk * TILE_M + (int)pair_rank * GEMM_BM + quad * 32;
            const int row = row0 + lane;
            const bool row_ok = row < p.M;
            const int col0 = n_blk * BN + c * 32;
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + as * Cfg::ACC_COLS + c * 32 + ((uint32_t)(quad * 32) << 16), v);
            // stage this chunk's bias while the TMEM load is in flight
            __syncwarp();
            const bool gelu_half = OUT_BF16 && p.act == 1 && p.out_f16;   // GELU done below in packed fp16, on x / 2
            const float pre_scale = gelu_half ? 0.5f : 1.0f;
            bias_w[lane] = (p.bias != nullptr && col0 + lane < p.N) ? pre_scale * __ldg(p.bias + col0 + lane) : 0.0f;
            __syncwarp();
            tmem_ld_wait();
            if (nc == grp) {                       // last chunk of the tile: all TMEM reads of this tile by this warp are done
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (PAIR) mbar_arrive_leader(&tempty_bar[as]);
                    else mbar_arrive(&tempty_bar[as]);
                }
            }
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_w + j);
                f[j + 0] = fmaf(__uint_as_float(v[j + 0]), pre_scale, b4.x);
                f[j + 1] = fmaf(__uint_as_float(v[j + 1]), pre_scale, b4.y);
                f[j + 2] = fmaf(__uint_as_float(v[j + 2]), pre_scale, b4.z);
                f[j + 3] = fmaf(__uint_as_float(v[j + 3]), pre_scale, b4.w);
            }
            if (p.act == 1 && !gelu_half) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
            } else if (p.act == 2) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
            }
            uint8_t* sbuf = cst + b * Cfg::CST;
            if constexpr (MUL) {
                {
                    mbar_wait(&rbar[b], (i / NBUF) & 1);
                    const uint8_t* rowp = sbuf + lane * 64;
                    const int sw = (lane >> 1) & 3;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 u = *reinterpret_cast<const uint4*>(rowp + ((q ^ sw) << 4));
                        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 gd = gelu_erf_grad_bf16x2(w4[k]);   // packed fp16 evaluation, fp32 product (gradients underflow fp16)
                            f[q * 8 + 2 * k] *= gd.x;
                            f[q * 8 + 2 * k + 1] *= gd.y;
                        }
                    }
                }
            }
            if constexpr (!OUT_BF16) {
                if (p.aux != nullptr && row_ok) {
                    float* ap = p.aux + ((long long)(row / p.aux_T) * p.aux_bstride + (row % p.aux_T)) * p.ld_aux + col0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (col0 + j < p.N) *reinterpret_cast<float4*>(ap + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                }
                if (tma_resid) {   // (never MUL here: MUL implies a 16-bit output)
                    mbar_wait(&rbar[b], (i / NBUF) & 1);
                    const uint8_t* rowp = sbuf + lane * 128;
                    const int sw = lane & 7;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 r4 = *reinterpret_cast<const float4*>(rowp + ((q ^ sw) << 4));
                        f[q * 4] += r4.x; f[q * 4 + 1] += r4.y; f[q * 4 + 2] += r4.z; f[q * 4 + 3] += r4.w;
                    }
                }
                if (p.resid2 != nullptr && row_ok) {
                    const float* rp = p.resid2 + (long long)row * p.ldr2 + col0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        if (col0 + j < p.N) {
                            const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
                            f[j] += r4.x; f[j + 1] += r4.y; f[j + 2] += r4.z; f[j + 3] += r4.w;
                        }
                    }
                }
            }
            if (!tma_resid) {                      // buffer b was last handed to TMA NBUF chunks ago: make sure it has been read out
                if (lane == 0) tma_store_wait_read<NBUF - 1>();
                __syncwarp();
            }
            if constexpr (OUT_BF16) {
                // row = 64 B (32 x 16-bit), CU_TENSOR_MAP_SWIZZLE_64B: 16-byte unit index ^= (row >> 1) & 3
                uint8_t* rowp = sbuf + lane * 64;
                const int sw = (lane >> 1) & 3;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 u;
                    if (gelu_half) {
                        u.x = gelu_erf_f16x2_halved(f[q * 8 + 0], f[q * 8 + 1]);
                        u.y = gelu_erf_f16x2_halved(f[q * 8 + 2], f[q * 8 + 3]);
                        u.z = gelu_erf_f16x2_halved(f[q * 8 + 4], f[q * 8 + 5]);
                        u.w = gelu_erf_f16x2_halved(f[q * 8 + 6], f[q * 8 + 7]);
                    } else {
                        u.x = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
                        u.y = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
                        u.z = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
                        u.w = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
                    }
                    *reinterpret_cast<uint4*>(rowp + ((q ^ sw) << 4)) = u;
                }
            } else {
                // row = 128 B (32 fp32), CU_TENSOR_MAP_SWIZZLE_128B: 16-byte unit index ^= row & 7
                uint8_t* rowp = sbuf + lane * 128;
                const int sw = lane & 7;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(rowp + ((q ^ sw) << 4)) = make_float4(f[q * 4 + 0], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&tmC, sbuf, col0, row0 + ks * p.split_rows);
                tma_store_commit();
            }
            ++i;
            if (nc == grp) ++it;
            tile = ntile;
            c = nc;
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();   // the leader's MMAs read the peer's shared memory: nobody leaves early
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled_local g_encode = nullptr;

static int ensure_encode() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) return set_error(ARD_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_local>(fn);
    return 0;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
    if (int rc = ensure_encode()) return rc;
    cuuint64_t gdim[2] = {inner, outer};
    cuuint64_t gstride[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = g_encode(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(ARD_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) base=%p inner=%llu outer=%llu stride=%llu", (int)r,
                                            base, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes);
    return 0;
}

template <int BN, bool OUT_BF16, bool PAIR, bool MUL = false, int NEPI = 8>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tr, const GemmKernelParams& kp,
                       int num_sms, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, OUT_BF16, PAIR, MUL, NEPI>;
    static_assert(Cfg::SMEM_BYTES <= 227 * 1024 && Cfg::STAGES >= 3, "gemm: shared memory budget");
    auto kern = gemm_tc_kernel<BN, OUT_BF16, PAIR, MUL, NEPI>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return set_error(ARD_ERR_CUDA, "cudaFuncSetAttribute(gemm smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
        attr_set = true;
    }
    constexpr int TILE_M = PAIR ? 2 * GEMM_BM : GEMM_BM;
    const int tiles = ((kp.M + TILE_M - 1) / TILE_M) * ((kp.N + BN - 1) / BN) * kp.splitk;
    if constexpr (PAIR) {
        // one CTA pair (cluster of 2, same TPC) per 256 x BN tile
        const int pairs = tiles < num_sms / 2 ? tiles : num_sms / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(Cfg::THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = pdl_enabled() ? 2 : 1;
        return check_cuda(cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, tr, kp), "gemm pair launch");
    } else {
        const int grid = tiles < num_sms ? tiles : num_sms;
        ARD_CUDA(enqueue_pdl(kern, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, ta, tb, tc, tr, kp));
        return check_cuda(cudaGetLastError(), "gemm launch");
    }
}

// CTA pairs (cta_group::2) halve the W-tile bytes each SM stages and reads per MMA. Measured on B200 (tools/bench_ops.py,
// profiles/r1_gemm_pair.md): a gain for the deep-K fp32-output GEMMs (fc2 of stages 2/3, K >= 1536: 103 -> 95 us, 91 -> 84 us),
// a loss for the K = 384 16-bit-output ones (the two CTAs' epilogues gate one shared accumulator hand-off), so the
// heuristic only pairs the former. ARD_GEMM_PAIR=0/1 forces never/always (read per call so tests can toggle it).
static bool pick_pair(const GemmArgs& a, int BN) {
    const char* e = getenv("ARD_GEMM_PAIR");
    const int mode = e ? (atoi(e) ? 1 : 0) : 2;
    if (a.force_pair < 0 || mode == 0 || BN < 128) return false;
    if (a.force_pair > 0 || mode == 1) return true;
    return !a.out_bf16 && a.K >= 1536 && a.M >= 2048;
}

static int pick_bn(int N, bool out16) {
    // widest tile that divides N (fewest re-reads of A); fall back to 128 with a clipped last tile.
    // fp32-output kernels spend their shared memory on 3 epilogue staging buffers per warp and stop at BN = 192.
    if (const char* e = getenv("ARD_GEMM_BN")) {   // development override of the N tile
        const int forced = atoi(e);
        if (forced && (out16 || forced <= 192)) return forced;
    }
    if (out16 && N % 256 == 0) return 256;
    if (N % 192 == 0) return 192;
    if (N % 128 == 0) return 128;
    if (N % 96 == 0) return 96;
    return 128;
}

int gemm_bf16(const GemmArgs& a, int num_sms, cudaStream_t stream) {
    if (a.M <= 0 || a.N <= 0 || a.K <= 0 || (a.K % 16) != 0 || (a.lda % 8) != 0 || (a.ldw % 8) != 0)
        return set_error(ARD_ERR_SHAPE, "gemm: bad shape M=%d N=%d K=%d lda=%lld ldw=%lld", a.M, a.N, a.K, a.lda, a.ldw);
    if ((a.out_bf16 && (a.ldo % 8) != 0) || (!a.out_bf16 && (a.ldo % 4) != 0)) return set_error(ARD_ERR_SHAPE, "gemm: ldo alignment");
    if (a.out_bf16 && (a.resid1 || a.resid2 || a.aux)) return set_error(ARD_ERR_SHAPE, "gemm: residual/aux need fp32 output");
    if (a.out_f16 && !(a.out_bf16 && a.act == ARD_ACT_GELU)) return set_error(ARD_ERR_SHAPE, "gemm: fp16 output is only produced by the GELU epilogue");
    int BN = a.force_bn ? a.force_bn : pick_bn(a.N, a.out_bf16 != 0);
    if (a.mul_gelu_bwd != nullptr) BN = (a.N % 192 == 0) ? 192 : 128;   // the multiplicand kernels are built for these two tiles
    if (!a.out_bf16 && BN > 192) return set_error(ARD_ERR_SHAPE, "gemm: fp32 output supports BN <= 192");
    CUtensorMap ta, tb, tc, tr;
    if (int rc = make_tmap_2d(&ta, a.A, 2, a.K, a.M, (uint64_t)a.lda * 2, GEMM_BK, GEMM_BM, 128)) return rc;
    const bool pair = a.mul_gelu_bwd == nullptr && pick_pair(a, BN);
    if (int rc = make_tmap_2d(&tb, a.W, 2, a.K, a.N, (uint64_t)a.ldw * 2, GEMM_BK, pair ? BN / 2 : BN, 128)) return rc;
    const int splitk = a.splitk > 1 ? a.splitk : 1;
    if (splitk > 1) {
        if (a.out_bf16 || a.bias || a.act || a.resid1 || a.resid2 || a.aux || a.mul_gelu_bwd || a.split_rows < a.M || (a.split_rows % 256) != 0)
            return set_error(ARD_ERR_SHAPE, "gemm: split-K needs a plain fp32 output and split_rows (a multiple of 256) >= M");
    }
    if (a.out_bf16) {
        if (int rc = make_tmap_2d(&tc, a.out, 2, a.N, a.M, (uint64_t)a.ldo * 2, 32, 32, 64)) return rc;
    } else {
        const uint64_t out_rows = splitk > 1 ? (uint64_t)splitk * a.split_rows : (uint64_t)a.M;
        if (int rc = make_tmap_2d(&tc, a.out, 4, a.N, out_rows, (uint64_t)a.ldo * 4, 32, 32, 128)) return rc;
    }
    tr = tc;
    if (!a.out_bf16 && a.resid1 != nullptr) {
        if (a.ldr1 % 4) return set_error(ARD_ERR_SHAPE, "gemm: residual leading dimension must be a multiple of 4");
        if (int rc = make_tmap_2d(&tr, a.resid1, 4, a.N, a.M, (uint64_t)a.ldr1 * 4, 32, 32, 128)) return rc;
    }
    if (a.mul_gelu_bwd != nullptr) {
        if (!a.out_bf16 || a.out_f16 || a.act != ARD_ACT_NONE || (a.ld_mul % 8))
            return set_error(ARD_ERR_SHAPE, "gemm: the gelu' multiplicand needs a plain bf16 output and ld_mul %% 8 == 0");
        if (int rc = make_tmap_2d(&tr, a.mul_gelu_bwd, 2, a.N, a.M, (uint64_t)a.ld_mul * 2, 32, 32, 64)) return rc;
    }
    GemmKernelParams kp;
    kp.M = a.M; kp.N = a.N; kp.K = a.K;
    kp.bias = a.bias; kp.act = a.act;
    kp.resid1 = a.resid1; kp.ldr1 = a.ldr1; kp.resid2 = a.resid2; kp.ldr2 = a.ldr2;
    kp.aux = a.aux; kp.ld_aux = a.ld_aux; kp.aux_T = a.aux_T > 0 ? a.aux_T : a.M; kp.aux_bstride = a.aux_bstride;
    kp.ab_f16 = a.ab_f16; kp.out_f16 = a.out_f16; kp.mul_gelu_bwd = a.mul_gelu_bwd != nullptr;
    {
        const int num_kb = (a.K + GEMM_BK - 1) / GEMM_BK;
        kp.splitk = splitk;
        kp.kb_per_split = (num_kb + splitk - 1) / splitk;
        kp.splitk = (num_kb + kp.kb_per_split - 1) / kp.kb_per_split;   // no empty splits (an empty k-range would never commit its accumulator)
        kp.split_rows = splitk > 1 ? a.split_rows : 0;
    }
    const double osz = a.out_bf16 ? 2.0 : 4.0;
    ProfScope ps(PROF_GEMM, stream, 2.0 * a.M * a.N * a.K,
                 2.0 * a.M * a.K + 2.0 * a.N * a.K + osz * a.M * a.N + (a.resid1 ? 4.0 * a.M * a.N : 0.0) + (a.resid2 ? 4.0 * a.M * a.N : 0.0) +
                     (a.aux ? 4.0 * a.M * a.N : 0.0));
#define ARD_GEMM_CASE(bn)                                                                                          \
    case bn:                                                                                                       \
        return a.out_bf16 ? launch_gemm<bn, true, false>(ta, tb, tc, tr, kp, num_sms, stream)                      \
                          : launch_gemm<bn, false, false>(ta, tb, tc, tr, kp, num_sms, stream);
#define ARD_GEMM_PAIR_CASE(bn)                                                                                     \
    case bn:                                                                                                       \
        return a.out_bf16 ? launch_gemm<bn, true, true>(ta, tb, tc, tr, kp, num_sms, stream)                       \
                          : launch_gemm<bn, false, true>(ta, tb, tc, tr, kp, num_sms, stream);
    if (a.mul_gelu_bwd != nullptr) {
        if (BN == 192) return launch_gemm<192, true, false, true, 16>(ta, tb, tc, tr, kp, num_sms, stream);
        return launch_gemm<128, true, false, true, 16>(ta, tb, tc, tr, kp, num_sms, stream);
    }
    if (a.out_bf16 && a.act == ARD_ACT_GELU && !pair && a.K <= 384) {   // epilogue-bound (short K loop): 16 epilogue warps
        switch (BN) {
            case 96: return launch_gemm<96, true, false, false, 16>(ta, tb, tc, tr, kp, num_sms, stream);
            case 128: return launch_gemm<128, true, false, false, 16>(ta, tb, tc, tr, kp, num_sms, stream);
            case 192: return launch_gemm<192, true, false, false, 16>(ta, tb, tc, tr, kp, num_sms, stream);
            case 256: return launch_gemm<256, true, false, false, 16>(ta, tb, tc, tr, kp, num_sms, stream);
        }
    }
    if (pair) {
        switch (BN) {
            ARD_GEMM_PAIR_CASE(128)
            ARD_GEMM_PAIR_CASE(192)
            case 256:
                if (a.out_bf16) return launch_gemm<256, true, true>(ta, tb, tc, tr, kp, num_sms, stream);
                break;
        }
        return set_error(ARD_ERR_SHAPE, "gemm: unsupported pair BN=%d", BN);
    }
    switch (BN) {
        ARD_GEMM_CASE(96)
        ARD_GEMM_CASE(128)
        ARD_GEMM_CASE(192)
        case 256:
            if (a.out_bf16) return launch_gemm<256, true, false>(ta, tb, tc, tr, kp, num_sms, stream);
            break;
    }
#undef ARD_GEMM_PAIR_CASE
#undef ARD_GEMM_CASE
    return set_error(ARD_ERR_SHAPE, "gemm: unsupported BN=%d", BN);
}

}  // namespace ard
