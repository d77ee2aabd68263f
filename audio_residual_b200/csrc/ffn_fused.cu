// Fused Swin FFN for the 96-channel stage:  out = x + fc2(GELU(fc1(LayerNorm(x))))  (+ optional second residual)
//
// Reference: SwinTransformerBlock.forward htsat.py:479-480 (`x = x + drop_path(mlp(norm2(x)))`, Mlp htsat.py:158-164) and
// the doubled form of the ResiDual-patched block, src/residual.py:93-96.
//
// Unfused, one FFN at stage 0 moves 34 bytes per token-channel through HBM (LN out, 4C hidden written and re-read) and is
// ~6x over its HBM floor; here x is read once (+ once more for the residual add, from L2) and the result written once.
// Both weight matrices (2 x 72 KB) stay resident in shared memory for the life of the persistent CTA; the 4C-wide hidden
// activation only ever exists in TENSOR MEMORY: as a 128x64 fp32 accumulator, then as the packed-fp16 A operand of fc2.
//
// Persistent CTA (one per SM), 26 warps, every role loops over the CTA's 128-token tiles:
//   warps 0-7   : output epilogue (TMEM lane quadrant w & 3, column half w >> 2): Y accumulator -> + b2 + x (+ resid2) ->
//                 swizzled staging -> TMA store; the residual tiles are prefetched by TMA one 16-column chunk ahead
//   warp 8      : TMEM allocator, one-time TMA load of W1/W2, fc1 issue (warp-convergent loop, elected lane). Per PAIR of 64-wide
//                 hidden chunks: H = A1 W1^T (M128 N128 K96) into two of the four H accumulators, so the tensor core runs up
//                 to two pairs ahead of the GELU warps
//   warps 9-24  : 16 GELU warps in two groups on alternate chunks (4 per TMEM lane quadrant, 32 hidden columns each):
//                 tcgen05.ld H_j, + b1, erf GELU in packed fp16 math, tcgen05.st of the fp16 result into the A2 columns of
//                 tensor memory. The same warps LayerNorm the NEXT tile (8 rows per warp, 8 lanes per row, 3-step butterflies,
//                 bf16 rows written straight into the SWIZZLE_64B K-major A-operand layout), group 0 after its first chunk
//                 of the current tile and group 1 after its second
//   warp 25     : fc2 issue: Y += A2_j W2[:, j]^T (M128 N96 K64, fp16 x fp16) with the A operand read from tensor memory
//
// Measured history of this kernel is in profiles/r1_ffn_fused.md (520 -> 251 us at M = 1,048,576): the per-role clock trace
// (tools/ffn_trace.py) showed a dedicated LayerNorm role serialising with the fc1 issue train behind a single A1 buffer; the
// shared-memory hand-off of the GELU output cost an async-proxy fence per chunk; row-per-thread residual loads in the epilogue
// cost 160 us with a second residual. ncu now: issue-active 63-66 %, DRAM traffic = algorithmic.
#include "ard_common.cuh"
#include "ard_internal.h"

namespace ard {

// Development aid (tools/ffn_trace.py builds a separate library with -DARD_FFN_TRACE): per-role clock64() stamps of CTA 0.
#ifdef ARD_FFN_TRACE
__device__ long long g_ffn_trace[8][64][8];   // [role][event index][field]
#define FF_TRACE(role, idx, field) do { if (blockIdx.x == 0 && (idx) < 64) g_ffn_trace[role][idx][field] = clock64(); } while (0)
#else
#define FF_TRACE(role, idx, field) do { } while (0)
#endif

constexpr int FF_C = 96;
constexpr int FF_HD = 4 * FF_C;         // 384
constexpr int FF_NCH = FF_HD / 64;      // 6 hidden chunks of 64
constexpr int FF_NPAIR = FF_NCH / 2;    // fc1 is issued per PAIR of chunks (N = 128 per MMA)
constexpr int FF_BM = 128;
constexpr int FF_EPI_WARPS = 8;
constexpr int FF_GELU_WARPS = 16;
constexpr int FF_W_MMA = FF_EPI_WARPS, FF_W_GELU = FF_W_MMA + 1, FF_W_MMA2 = FF_W_GELU + FF_GELU_WARPS;   // fc1 issuer, GELU+LN, fc2 issuer
constexpr int FF_THREADS = (FF_W_MMA2 + 1) * 32;   // 832

constexpr int FF_W1_KB = FF_HD * 64;            // bytes per 32-wide k-block of W1 (384 rows x 64 B)
constexpr int FF_W1_OFF = 0;
constexpr int FF_W1_BYTES = 3 * FF_W1_KB;       // 73728
constexpr int FF_W2_KB = FF_C * 128;            // bytes per 64-wide k-block of W2 (96 rows x 128 B)
constexpr int FF_W2_OFF = FF_W1_OFF + FF_W1_BYTES;
constexpr int FF_W2_BYTES = FF_NCH * FF_W2_KB;  // 73728
constexpr int FF_A1_KB = FF_BM * 64;            // 8192
constexpr int FF_A1_OFF = FF_W2_OFF + FF_W2_BYTES;
constexpr int FF_A1_BYTES = 3 * FF_A1_KB;       // 24576
constexpr int FF_A2_OFF = FF_A1_OFF + FF_A1_BYTES;
constexpr int FF_A2_BYTES = FF_BM * 128;        // the GELU output goes to fc2 through tensor memory; the 2 x 16 KB a shared-memory A2
                                                // tile would need hold the TMA-prefetched residual tiles: 8 warps x {x, resid2} x (32 rows x 64 B)
constexpr int FF_RB_OFF = FF_A2_OFF;
constexpr int FF_CST_OFF = FF_A2_OFF + 2 * FF_A2_BYTES;    // 8 warps x (32 rows x 64 B)
constexpr int FF_VEC_OFF = FF_CST_OFF + FF_EPI_WARPS * 2048;   // b1[384] b2[96] gamma[96] beta[96]
constexpr int FF_BAR_OFF = FF_VEC_OFF + (FF_HD + 3 * FF_C) * 4;
constexpr int FF_SMEM_BYTES = FF_BAR_OFF + 256 + 1024;

constexpr int FF_NHB = 4;       // H accumulators in flight (two pairs)
constexpr int FF_TM_H = 0;      // TMEM columns: H0..H3 @0/64/128/192, Y0 @256 (96 used), Y1 @384
constexpr int FF_TM_Y = 256;
constexpr int FF_TM_A2 = 352;   // A2 buffer g: 32 columns (64 fp16 per row) at 352 + 128 g, in the gaps after Y0 / Y1

struct FfnParams {
    const float* x;        // [M, 96] fp32  (LayerNorm input and first residual)
    const float* resid2;   // [M, 96] fp32 or null
    const float* gamma;    // norm2
    const float* beta;
    const float* b1;       // [384]
    const float* b2;       // [96]
    int M;
};

__global__ void __launch_bounds__(FF_THREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmX,
                 const __grid_constant__ CUtensorMap tmR2, const FfnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, still a shared-space pointer
    float* b1s = reinterpret_cast<float*>(smem + FF_VEC_OFF);
    float* b2s = b1s + FF_HD;
    float* gs = b2s + FF_C;
    float* bs = gs + FF_C;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FF_BAR_OFF);
    uint64_t* w_full = bars + 0;
    uint64_t* a1_full = bars + 1;
    uint64_t* a1_free = bars + 2;
    uint64_t* h_full = bars + 3;    // [4]
    uint64_t* h_free = bars + 7;    // [4]
    uint64_t* a2_full = bars + 11;  // [2]
    uint64_t* a2_free = bars + 13;  // [2]
    uint64_t* y_full = bars + 15;   // [2]
    uint64_t* y_free = bars + 17;   // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 19);
    uint64_t* rbar = bars + 20;     // [8] residual tiles landed (one per epilogue warp)

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = (p.M + FF_BM - 1) / FF_BM;
    const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    // (biases / LayerNorm affine are weights, never written by a predecessor kernel: safe to read before pdl_wait)
    for (int i = threadIdx.x; i < FF_HD; i += FF_THREADS) b1s[i] = 0.5f * p.b1[i];   // the GELU takes x / 2 (gelu_erf_f16x2_halved)
    for (int i = threadIdx.x; i < FF_C; i += FF_THREADS) {
        b2s[i] = p.b2[i];
        gs[i] = p.gamma[i];
        bs[i] = p.beta[i];
    }
    if (warp == FF_W_MMA && lane == 0) {
        tma_prefetch_desc(&tmW1);
        tma_prefetch_desc(&tmW2);
        tma_prefetch_desc(&tmOut);
        mbar_init(w_full, 1);
        mbar_init(a1_full, FF_GELU_WARPS);
        mbar_init(a1_free, 1);
        for (int i = 0; i < FF_NHB; ++i) {
            mbar_init(&h_full[i], 1);
            mbar_init(&h_free[i], FF_GELU_WARPS / 2);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a2_full[i], FF_GELU_WARPS / 2);
            mbar_init(&a2_free[i], 1);
            mbar_init(&y_full[i], 1);
            mbar_init(&y_free[i], FF_EPI_WARPS);
        }
        for (int i = 0; i < FF_EPI_WARPS; ++i) mbar_init(&rbar[i], 1);
        fence_barrier_init();
    }
    if (warp == FF_W_MMA) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();

    if (warp < FF_EPI_WARPS) {
        // ============================================================ output epilogue warps
        // warp w: TMEM lane quadrant w & 3 (rows 32(w&3)..), column half w >> 2 (three 16-column chunks)
        const int quad = warp & 3, c_begin = (warp >> 2) * 3;
        uint8_t* sbuf = smem + FF_CST_OFF + warp * 2048;
        const bool has_r2 = p.resid2 != nullptr;
        // The residual tiles (x, and the second residual) of a chunk are fetched by TMA one chunk ahead into this warp's two
        // 2 KB buffers: coalesced, asynchronous, no registers held across the wait. (Row-per-thread LDG.128 touches 32 lines
        // per instruction; with a second residual it cost 160 us per launch.) Chunk sequence number q = 3 * it + cc.
        uint8_t* rb1 = smem + FF_RB_OFF + warp * 4096;
        uint8_t* rb2 = rb1 + 2048;
        auto fetch_resid = [&](int tile, int c) {       // lane 0 only
            mbar_expect_tx(&rbar[warp], has_r2 ? 4096 : 2048);
            tma_load_2d(rb1, &tmX, &rbar[warp], c * 16, tile * FF_BM + quad * 32);
            if (has_r2) tma_load_2d(rb2, &tmR2, &rbar[warp], c * 16, tile * FF_BM + quad * 32);
        };
        if (lane == 0 && (int)blockIdx.x < num_tiles) fetch_resid(blockIdx.x, c_begin);
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int yb = it & 1;
            if (warp == 0 && lane == 0) FF_TRACE(0, it, 0);
            mbar_wait_parked(&y_full[yb], (it >> 1) & 1);
            if (warp == 0 && lane == 0) FF_TRACE(0, it, 1);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < 3; ++cc) {                 // 16-column chunks
                const int c = c_begin + cc;
                uint32_t v[16];
                tmem_ld_32x32b_x16(tmem_base + FF_TM_Y + yb * 128 + c * 16 + ((uint32_t)(quad * 32) << 16), v);
                mbar_wait(&rbar[warp], (uint32_t)((it * 3 + cc) & 1));
                const int sw = (lane >> 1) & 3;              // SWIZZLE_64B: 16-byte unit index ^= (row >> 1) & 3
                float4 r1[4], r2[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    r1[j] = *reinterpret_cast<const float4*>(rb1 + lane * 64 + ((j ^ sw) << 4));
                    r2[j] = has_r2 ? *reinterpret_cast<const float4*>(rb2 + lane * 64 + ((j ^ sw) << 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                fence_proxy_async_smem();                    // order this lane's generic-proxy reads before the async-proxy (TMA) overwrite
                __syncwarp();                                // every lane has read the buffers: the next fetch may overwrite them
                if (lane == 0) {
                    if (cc + 1 < 3) fetch_resid(tile, c + 1);
                    else if (tile + (int)gridDim.x < num_tiles) fetch_resid(tile + gridDim.x, c_begin);
                }
                tmem_ld_wait();
                if (cc == 2) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&y_free[yb]);
                }
                float4 o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 b4 = *reinterpret_cast<const float4*>(b2s + c * 16 + j * 4);
                    o[j].x = __uint_as_float(v[j * 4 + 0]) + b4.x + r1[j].x + r2[j].x;
                    o[j].y = __uint_as_float(v[j * 4 + 1]) + b4.y + r1[j].y + r2[j].y;
                    o[j].z = __uint_as_float(v[j * 4 + 2]) + b4.z + r1[j].z + r2[j].z;
                    o[j].w = __uint_as_float(v[j * 4 + 3]) + b4.w + r1[j].w + r2[j].w;
                }
                if (lane == 0) tma_store_wait_read<0>();     // the previous chunk's store has read the (single) staging buffer
                __syncwarp();
                uint8_t* rowp = sbuf + lane * 64;
#pragma unroll
                for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(rowp + ((j ^ sw) << 4)) = o[j];
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmOut, sbuf, c * 16, tile * FF_BM + quad * 32);
                    tma_store_commit();
                }
            }
            if (warp == 0 && lane == 0) FF_TRACE(0, it, 2);
        }
        if (lane == 0) tma_store_wait_all<0>();
    } else if (warp == FF_W_MMA) {
        // ============================================================ weight load + fc1 MMA issue
        // The whole warp runs the loop (all lanes wait on the barriers) and one elected lane issues: in warp-convergent code
        // ptxas keeps descriptors / TMEM addresses / loop state on the uniform datapath and emits the UTCHMMAs back to back
        // (issued from a lane-0 branch each MMA cost ~12 SASS instructions: R2UR per operand + an ELECT / BRA.U.ANY wrapper).
        if (elect_one_sync()) {
            mbar_expect_tx(w_full, FF_W1_BYTES + FF_W2_BYTES);
            for (int kb = 0; kb < 3; ++kb)
                for (int hf = 0; hf < 2; ++hf)
                    tma_load_2d(smem + FF_W1_OFF + kb * FF_W1_KB + hf * 192 * 64, &tmW1, w_full, kb * 32, hf * 192);
            for (int j = 0; j < FF_NCH; ++j) tma_load_2d(smem + FF_W2_OFF + j * FF_W2_KB, &tmW2, w_full, j * 64, 0);
        }
        __syncwarp();
        mbar_wait_parked(w_full, 0);
        // fc1 per PAIR of hidden chunks: H[:, 128 jp ..] = A1 W1[128 jp .., :]^T as M128 N128 K16 MMAs (N = 64 MMAs read 6 KB of
        // operands per 32 tensor-cycles and ran at half rate on shared-memory bandwidth; N = 128 reads 8 KB per 64). Pair
        // P = 3 t + jp lands in H buffers 2 (P & 1), 2 (P & 1) + 1; the issuer runs up to two pairs ahead of the GELU warps
        // (bounded by h_free) and across tile boundaries (bounded by a1_full). fc2 is issued by a second warp (FF_W_MMA2).
        constexpr uint32_t idesc1 = umma_idesc_bf16(FF_BM, 128);
        const uint64_t dA1 = umma_desc_sw64(smem_u32(smem + FF_A1_OFF)), dW1 = umma_desc_sw64(smem_u32(smem + FF_W1_OFF));
        int P = 0;
        for (int t = 0; t < my_tiles; ++t) {
            if (lane == 0) FF_TRACE(1, P, 0);
            mbar_wait_parked(a1_full, t & 1);                       // this tile's LayerNorm output is in A1
#pragma unroll 1
            for (int jp = 0; jp < FF_NPAIR; ++jp, ++P) {
                const int hb = (P & 1) * 2;
                const uint32_t par = ((P >> 1) & 1) ^ 1;     // buffers hb, hb+1 are on their (P >> 1)-th use
                if (lane == 0) FF_TRACE(1, P, 1);
                mbar_wait_parked(&h_free[hb], par);
                mbar_wait_parked(&h_free[hb + 1], par);
                if (lane == 0) FF_TRACE(1, P, 2);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t d = tmem_base + FF_TM_H + hb * 64;
                    const uint64_t db = dW1 + (uint64_t)(jp * ((128 * 64) >> 4));   // descriptor start address is in 16-byte units
#pragma unroll
                    for (int kb = 0; kb < 3; ++kb)
                        umma_f16_ss_run<2>(d, dA1 + (uint64_t)(kb * (FF_A1_KB >> 4)), db + (uint64_t)(kb * (FF_W1_KB >> 4)), idesc1, kb != 0);
                    umma_commit(&h_full[hb]);
                    umma_commit(&h_full[hb + 1]);
                    if (jp == FF_NPAIR - 1) umma_commit(a1_free);
                }
                __syncwarp();
                if (lane == 0) FF_TRACE(1, P, 3);
            }
        }
    } else if (warp == FF_W_MMA2) {
        // ============================================================ fc2 issuer: Y += A2_g W2[:, j]^T (same convergent pattern)
        mbar_wait_parked(w_full, 0);
        constexpr uint32_t idesc2 = umma_idesc_f16(FF_BM, FF_C);   // A2 (GELU output) and W2 are fp16
        const uint64_t dW2 = umma_desc_sw128(smem_u32(smem + FF_W2_OFF));
        int g = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int yb = t & 1;
#pragma unroll 1
            for (int j = 0; j < FF_NCH; ++j, ++g) {
                const int b = g & 1;
                if (lane == 0) FF_TRACE(2, g, 0);
                mbar_wait_parked(&a2_full[b], (g >> 1) & 1);
                if (lane == 0) FF_TRACE(2, g, 1);
                if (j == 0) mbar_wait_parked(&y_free[yb], ((t >> 1) & 1) ^ 1);
                if (lane == 0) FF_TRACE(2, g, 2);
                tc_fence_after();
                if (elect_one_sync()) {
                    umma_f16_ts_run4(tmem_base + FF_TM_Y + yb * 128, tmem_base + FF_TM_A2 + b * 128, dW2 + (uint64_t)(j * (FF_W2_KB >> 4)), idesc2,
                                     j != 0);
                    umma_commit(&a2_free[b]);
                    if (j == FF_NCH - 1) umma_commit(&y_full[yb]);
                }
                __syncwarp();
                if (lane == 0) FF_TRACE(2, g, 3);
            }
        }
    } else {
        // ============================================================ GELU + LayerNorm warps
        // 16 warps in two groups: group g turns the hidden chunks j = g, g+2, g+4 of every tile (TMEM, fp32) into the fp16 A2
        // operand of fc2 (+ b1, erf GELU in packed fp16 math, SWIZZLE_128B tile in shared memory), so that while one group sits
        // in the latency part of a chunk (barrier, TMEM load, async-proxy fence) the other is issuing math.
        // The same warps also produce the NEXT tile's LayerNorm operand A1: 8 rows per warp (8 lanes per row, 12 contiguous
        // channels each, 3-step butterflies, bf16 rows written straight into the SWIZZLE_64B K-major layout), group 0 after its
        // first chunk of the current tile and group 1 after its second. With four dedicated LayerNorm warps (32 rows each, one
        // latency-bound chain of ~1200 instructions) LN took 8.4 k cycles per tile and, A1 being single-buffered, serialised with
        // the 4 k-cycle fc1 issue train: 12.3 k cycles per tile while the GELU warps idled 42 % (tools/ffn_trace.py). Spread
        // over 16 warps it is ~1 k cycles in the shadow of the other group's GELU chunk, and A1 is free by then because fc1 of
        // the current tile is issued (pairwise) as soon as the first pair has been read out of TMEM.
        const int ew = warp - FF_W_GELU;
        const int quad = warp & 3;
        const int grp = ew >> 3;          // chunk parity this warp works on == A2 buffer it fills
        const int half = (ew >> 2) & 1;   // which 32 of the chunk's 64 hidden columns
        const int row = quad * 32 + lane; // row inside the tile == TMEM lane
        const int l8 = lane & 7, rsub = lane >> 3;     // LayerNorm mapping: 8 lanes per row, 4 rows per warp instruction
        uint8_t* a1 = smem + FF_A1_OFF;
        int a1_off[3];                                 // byte offset of this lane's three 8-byte stores for row-group 0
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int c0 = l8 * 12 + q * 4, kb = c0 >> 5, cc = c0 & 31;
            a1_off[q] = kb * FF_A1_KB + (ew * 8 + rsub) * 64 + (((cc >> 3) ^ (rsub >> 1)) << 4) + (cc & 7) * 2;
        }
        // pull a tile's x (and second-residual) rows of this warp into L2: 8 rows x 384 B = 24 lines of 128 B
        auto prefetch_tile = [&](int tile) {
            if (tile >= num_tiles || lane >= 24) return;
            const long long r0 = (long long)tile * FF_BM + ew * 8;
            if (r0 * FF_C * 4 + (lane + 1) * 128 > (long long)p.M * FF_C * 4) return;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.x + r0 * FF_C) + lane * 128));
            if (p.resid2 != nullptr)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.resid2 + r0 * FF_C) + lane * 128));
        };
        // LayerNorm of rows [ew*8, ew*8+8) of `tile` into A1; `lt` = this CTA's tile counter of `tile`
        auto layernorm_tile = [&](int tile, int lt) {
            const long long row0 = (long long)tile * FF_BM + ew * 8;
            float4 v[2][3];
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
                const long long r = row0 + gi * 4 + rsub;
                if (r < p.M) {
                    const float4* xr = reinterpret_cast<const float4*>(p.x + r * FF_C + l8 * 12);
                    v[gi][0] = __ldg(xr); v[gi][1] = __ldg(xr + 1); v[gi][2] = __ldg(xr + 2);
                } else {
                    v[gi][0] = v[gi][1] = v[gi][2] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            prefetch_tile(tile + gridDim.x);
            if (ew == 0 && lane == 0) FF_TRACE(3, lt, 0);
            mbar_wait_parked(a1_free, (lt & 1) ^ 1);                // fc1 MMAs of the previous tile have read A1
            if (ew == 0 && lane == 0) FF_TRACE(3, lt, 1);
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
                const float mean = sm[gi] * (1.0f / FF_C);
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
                const float rstd = rsqrtf(qv[gi] * (1.0f / FF_C) + 1e-5f);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int c0 = l8 * 12 + q * 4;                       // 4 channels = 8 bytes of bf16, inside one 16-byte unit
                    const float4 gm = *reinterpret_cast<const float4*>(gs + c0);
                    const float4 bt = *reinterpret_cast<const float4*>(bs + c0);
                    uint2 pk;
                    pk.x = pack_bf16x2(fmaf(v[gi][q].x * rstd, gm.x, bt.x), fmaf(v[gi][q].y * rstd, gm.y, bt.y));
                    pk.y = pack_bf16x2(fmaf(v[gi][q].z * rstd, gm.z, bt.z), fmaf(v[gi][q].w * rstd, gm.w, bt.w));
                    // row rr = ew*8 + gi*4 + rsub; SWIZZLE_64B unit index ^= (rr >> 1) & 3 = ((gi & 1) << 1) | (rsub >> 1): the
                    // gi-dependent part is a compile-time XOR of bit 5 plus a compile-time row offset
                    *reinterpret_cast<uint2*>(a1 + ((a1_off[q] ^ ((gi & 1) << 5)) + gi * 256)) = pk;
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a1_full);
            if (ew == 0 && lane == 0) FF_TRACE(3, lt, 2);
        };

        prefetch_tile(blockIdx.x);
        if (my_tiles > 0) layernorm_tile(blockIdx.x, 0);
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
#pragma unroll
            for (int jj = 0; jj < FF_NCH / 2; ++jj) {
                const int j = 2 * jj + grp;
                const int g = it * FF_NCH + j;              // global chunk index of this CTA
                const int hb = g & (FF_NHB - 1);
                if ((ew & 7) == 0 && lane == 0) FF_TRACE(4 + grp, g, 0);
                mbar_wait_parked(&h_full[hb], (g >> 2) & 1);
                if ((ew & 7) == 0 && lane == 0) FF_TRACE(4 + grp, g, 1);
                tc_fence_after();
                uint32_t pk[16];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(tmem_base + FF_TM_H + hb * 64 + half * 32 + s * 16 + ((uint32_t)(quad * 32) << 16), v);
                    tmem_ld_wait();
                    if (s == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&h_free[hb]);
                    }
                    const float* bb = b1s + j * 64 + half * 32 + s * 16;
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(bb + i);
                        pk[s * 8 + i / 2] = gelu_erf_f16x2_halved(fmaf(__uint_as_float(v[i]), 0.5f, b4.x), fmaf(__uint_as_float(v[i + 1]), 0.5f, b4.y));
                        pk[s * 8 + i / 2 + 1] = gelu_erf_f16x2_halved(fmaf(__uint_as_float(v[i + 2]), 0.5f, b4.z), fmaf(__uint_as_float(v[i + 3]), 0.5f, b4.w));
                    }
                }
                if ((ew & 7) == 0 && lane == 0) FF_TRACE(4 + grp, g, 2);
                mbar_wait_parked(&a2_free[grp], ((g >> 1) & 1) ^ 1);   // fc2 MMAs that read the previous contents have retired
                tc_fence_after();                               // order the tcgen05.st below after those MMAs' reads of the A2 columns
                if ((ew & 7) == 0 && lane == 0) FF_TRACE(4 + grp, g, 3);
                // 32 fp16 of this lane's row = 16 packed columns of the A2 tile in tensor memory (K pair 2c, 2c+1 in column c)
                tmem_st_32x32b_x16(tmem_base + FF_TM_A2 + grp * 128 + half * 16 + ((uint32_t)(quad * 32) << 16), pk);
                tmem_st_wait();
                tc_fence_before();
                (void)row;
                __syncwarp();
                if (lane == 0) mbar_arrive(&a2_full[grp]);
                if ((ew & 7) == 0 && lane == 0) FF_TRACE(4 + grp, g, 4);
                if (jj == grp && tile + (int)gridDim.x < num_tiles) layernorm_tile(tile + gridDim.x, it + 1);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == FF_W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// x_out = x + fc2(gelu(fc1(LN(x)))) (+ resid2). x_out may alias x (each element is read by the warp that later writes it).
int ffn_fused_96(const float* x, const float* resid2, float* out, long long M, const float* gamma, const float* beta,
                 const __nv_bfloat16* w1, const float* b1, const __half* w2_f16, const float* b2, int num_sms, cudaStream_t stream) {
    if (M <= 0) return 0;
    if (M > 0x7fffffffLL) return set_error(ARD_ERR_SHAPE, "ffn_fused: too many rows");
    CUtensorMap t1, t2, to;
    ARD_TRY(make_tmap_2d(&t1, w1, 2, FF_C, FF_HD, (uint64_t)FF_C * 2, 32, 192, 64));
    ARD_TRY(make_tmap_2d(&t2, w2_f16, 2, FF_HD, FF_C, (uint64_t)FF_HD * 2, 64, FF_C, 128));
    ARD_TRY(make_tmap_2d(&to, out, 4, FF_C, (uint64_t)M, (uint64_t)FF_C * 4, 16, 32, 64));
    CUtensorMap tx, tr;
    ARD_TRY(make_tmap_2d(&tx, x, 4, FF_C, (uint64_t)M, (uint64_t)FF_C * 4, 16, 32, 64));
    ARD_TRY(make_tmap_2d(&tr, resid2 ? resid2 : x, 4, FF_C, (uint64_t)M, (uint64_t)FF_C * 4, 16, 32, 64));
    static bool attr_set = false;
    if (!attr_set) {
        ARD_CUDA(cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM_BYTES));
        attr_set = true;
    }
    FfnParams p;
    p.x = x; p.resid2 = resid2; p.gamma = gamma; p.beta = beta; p.b1 = b1; p.b2 = b2; p.M = (int)M;
    const int tiles = (int)((M + FF_BM - 1) / FF_BM);
    const int grid = tiles < num_sms ? tiles : num_sms;
    const double MC = (double)M * FF_C;
    ProfScope ps(PROF_FFN, stream, 2.0 * M * FF_C * FF_HD * 2.0, MC * 4.0 * (2.0 + (resid2 ? 1.0 : 0.0)) + 2.0 * 2.0 * FF_C * FF_HD);
    ARD_CUDA(enqueue_pdl(ffn_fused_kernel, dim3(grid), dim3(FF_THREADS), FF_SMEM_BYTES, stream, t1, t2, to, tx, tr, p));
    return check_cuda(cudaGetLastError(), "ffn_fused launch");
}

}  // namespace ard

#ifdef ARD_FFN_TRACE
extern "C" int ard_debug_ffn_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, ard::g_ffn_trace, sizeof(long long) * 8 * 64 * 8);
}
#endif
