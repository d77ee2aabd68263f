// Fused Swin FFN for the 192- and 384-channel stages:  out = x + fc2(GELU(fc1(LayerNorm(x))))  (+ optional second residual)
//
// Reference: SwinTransformerBlock.forward htsat.py:479-480 (`x = x + drop_path(mlp(norm2(x)))`, Mlp htsat.py:158-164) and
// the doubled form of the ResiDual-patched block, src/residual.py:93-96.
//
// Same idea as ffn_fused.cu (C = 96): the LayerNorm output and the 4C-wide hidden activation never leave the SM. What
// changes with the width is that the weights (0.6 / 2.4 MB) no longer fit in shared memory: they are STREAMED from L2 through a
// ring of 24 KB slots by a TMA producer warp, once per 128-token tile, in the order the two MMA issuers consume them.
// Unfused, one FFN of these stages moves (LN out + hidden write + hidden read) 18 B per token-channel through HBM and
// re-reads the hidden activation once per output column tile from L2; here HBM sees x in and out only.
//
// Per 128-token tile and 64-wide hidden chunk j (NCH = 4C/64 chunks):
//   fc1:  H_j[128, 64]  = A1[128, C] W1[64j.., :]^T          (bf16, M128 N64 K16 MMAs, C/64 k-blocks, accumulator in TMEM)
//   GELU: A2_j[128, 64] = gelu(H_j + b1)                      (16 warps, packed fp16, written as the fp16 A operand of fc2)
//   fc2:  Y[128, C]    += A2_j W2[:, 64j..]^T                 (fp16, M128 N192 K16 MMAs, C/192 column blocks of Y)
//
// Warp roles (27 warps): 0-7 output epilogue (TMEM lane quadrant w & 3, column half w >> 2); 8 fc1 issuer (+ TMEM allocator);
// 9 TMA producer of the weight ring; 10-25 GELU warps in two groups on alternate chunks, which also LayerNorm the next
// tile's rows into A1 (8 rows per warp) once fc1 of the current tile has been issued; 26 fc2 issuer. Issuer and producer
// warps run warp-convergent loops with an elected lane (see gemm_tc.cu).
//
// Ring order per tile (one entry = C/192 slots): W1(0) W1(1) W2(0) W1(2) W2(1) ... W1(NCH-1) W2(NCH-3) W2(NCH-2) W2(NCH-1):
// fc1 runs two chunks (the two H accumulators) ahead of fc2, and every entry's consumption depends only on earlier entries.
#include "ard_common.cuh"
#include "ard_internal.h"

namespace ard {

// Development aid (tools/ffn_trace.py --wide builds a separate library with -DARD_FFN_TRACE): clock64() stamps of CTA 0.
#ifdef ARD_FFN_TRACE
__device__ long long g_ffw_trace[8][64][8];   // [role][event index][field]
#define FW_TRACE(role, idx, field) do { if (blockIdx.x == 0 && (idx) < 64) g_ffw_trace[role][idx][field] = clock64(); } while (0)
#else
#define FW_TRACE(role, idx, field) do { } while (0)
#endif

constexpr int FW_BM = 128;
constexpr int FW_EPI_WARPS = 8;
constexpr int FW_GELU_WARPS = 16;
constexpr int FW_W_MMA = FW_EPI_WARPS, FW_W_TMA = FW_W_MMA + 1, FW_W_GELU = FW_W_TMA + 1, FW_W_MMA2 = FW_W_GELU + FW_GELU_WARPS;
constexpr int FW_THREADS = (FW_W_MMA2 + 1) * 32;   // 864
constexpr int FW_TM_H = 384;                       // TMEM columns: Y (C, or 2 x 192) from 0, H0 / H1 at 384 / 448

template <int C>
struct FwCfg {
    static_assert(C == 128 || C == 192 || C == 256 || C == 384, "ffn_wide: C = 128, 192, 256 or 384");
    static constexpr int HD = 4 * C;
    static constexpr int NCH = HD / 64;              // hidden chunks per tile (8 / 12 / 16 / 24)
    // Ring slot = one YW-wide piece of the channel axis: KBS k-blocks of a W1 chunk [64 x 64 each] or one W2 block [YW x 64].
    // YW is also the N of the fc2 MMAs. 192 for the HTSAT-tiny widths (192, 384 = 2 x 192); the HTSAT-base widths 128 / 256 are one piece.
    static constexpr int YW = (C % 192 == 0) ? 192 : C;
    static constexpr int KBS = YW / 64;              // k-blocks per slot (2 / 3 / 4)
    static constexpr int SLOT = YW * 128;            // bytes: 16 / 24 / 32 KB
    static constexpr int NS = C / YW;                // slots per ring entry (2 at C = 384, else 1)
    static constexpr int KB1 = C / 64;               // fc1 k-blocks
    static constexpr int NYB = 1;                    // Y accumulators
    // C = 192: the GELU output reaches fc2 through TENSOR MEMORY (A operand from TMEM; Y 0-191, A2 192-255, H 384-511) and the
    // 32 KB of shared memory that held the A2 tiles hold the TMA-prefetched residual tiles of the epilogue instead (see
    // ffn_fused.cu: 372 -> 251 us there). At C = 384, Y + H fill all 512 TMEM columns: A2 and the residual loads stay as before.
    static constexpr bool A2T = C <= 256;            // Y 0..C-1, A2 C..C+63, H 384-511
    static constexpr int TM_A2 = C;
    // LayerNorm operand buffers / ring slots. Measured at C = 192 (B = 256, M = 262144): one A1 buffer + 5 slots 249 us, two A1
    // buffers (next tile's LayerNorm overlapped) + 3 slots 277 us: the ring depth is worth more than the overlap.
    static constexpr int NA1 = 1;
    static constexpr int NSLOT = C == 128 ? 8 : C == 192 ? 5 : 3;
    static constexpr int A1_KB = FW_BM * 128;        // 16384 bytes per 64-wide k-block of A1
    static constexpr int A1_OFF = 0;
    static constexpr int A1_BYTES = KB1 * A1_KB;
    static constexpr int A2_OFF = A1_OFF + NA1 * A1_BYTES;
    static constexpr int A2_BYTES = FW_BM * 128;     // 16384 per buffer
    static constexpr int RING_OFF = A2_OFF + 2 * A2_BYTES;
    static constexpr int CST_OFF = RING_OFF + NSLOT * SLOT;   // 8 warps x (32 rows x 64 B)
    static constexpr int VEC_OFF = CST_OFF + FW_EPI_WARPS * 2048;   // b2[C] gamma[C] beta[C]
    static constexpr int BAR_OFF = VEC_OFF + 3 * C * 4;
    static constexpr int SMEM_BYTES = BAR_OFF + 384 + 1024;
    static_assert(SMEM_BYTES <= 227 * 1024, "ffn_wide: shared memory budget");
    static constexpr int LN_CH = C / 16;             // channels per lane in the LayerNorm mapping (16 lanes per row)
    static constexpr int LN_V4 = LN_CH / 4;          // float4 per lane per row (3 / 6)
    static constexpr int LN_STEPS = LN_CH <= 12 ? 4 : 2;   // row-pairs held in registers at once (of the warp's 4): 32-48 fp32 registers
};

struct FwParams {
    const float* x;        // [M, C] fp32  (LayerNorm input and first residual)
    const float* resid2;   // [M, C] fp32 or null
    const float* gamma;    // norm2
    const float* beta;
    const float* b1h;      // [4C]  0.5 * fc1 bias (the packed GELU takes x / 2)
    const float* b2;       // [C]
    int M;
};

template <int C>
__global__ void __launch_bounds__(FW_THREADS, 1)
ffn_wide_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ CUtensorMap tmR2, const FwParams p) {
    using Cfg = FwCfg<C>;
    constexpr int NCH = Cfg::NCH, NS = Cfg::NS, NSLOT = Cfg::NSLOT, NYB = Cfg::NYB, NA1 = Cfg::NA1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, still a shared-space pointer
    float* b2s = reinterpret_cast<float*>(smem + Cfg::VEC_OFF);
    float* gs = b2s + C;
    float* bs = gs + C;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    uint64_t* ring_full = bars + 0;     // [8]
    uint64_t* ring_empty = bars + 8;    // [8]
    uint64_t* a1_full = bars + 16;      // [2]
    uint64_t* a1_free = bars + 18;      // [2]
    uint64_t* h_full = bars + 20;       // [2]
    uint64_t* h_free = bars + 22;       // [2]
    uint64_t* a2_full = bars + 24;      // [2]
    uint64_t* a2_free = bars + 26;      // [2]
    uint64_t* y_full = bars + 28;       // [2]
    uint64_t* y_free = bars + 30;       // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 32);
    uint64_t* rbar = bars + 34;         // [8] residual tiles landed (one per epilogue warp)

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = (p.M + FW_BM - 1) / FW_BM;
    const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    for (int i = threadIdx.x; i < C; i += FW_THREADS) {
        b2s[i] = p.b2[i];
        gs[i] = p.gamma[i];
        bs[i] = p.beta[i];
    }
    if (warp == FW_W_MMA && lane == 0) {
        tma_prefetch_desc(&tmW1);
        tma_prefetch_desc(&tmW2);
        tma_prefetch_desc(&tmOut);
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(&ring_full[i], 1);
            mbar_init(&ring_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a1_full[i], FW_GELU_WARPS);
            mbar_init(&a1_free[i], 1);
            mbar_init(&h_full[i], 1);
            mbar_init(&h_free[i], FW_GELU_WARPS / 2);
            mbar_init(&a2_full[i], FW_GELU_WARPS / 2);
            mbar_init(&a2_free[i], 1);
            mbar_init(&y_full[i], 1);
            mbar_init(&y_free[i], FW_EPI_WARPS);
        }
        for (int i = 0; i < FW_EPI_WARPS; ++i) mbar_init(&rbar[i], 1);
        fence_barrier_init();
    }
    if (warp == FW_W_MMA) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();

    if (warp < FW_EPI_WARPS) {
        // ============================================================ output epilogue warps
        // warp w: TMEM lane quadrant w & 3 (rows 32(w&3)..), column half w >> 2 (C/32 chunks of 16 columns)
        constexpr int NCC = C / 32;
        const int quad = warp & 3, c_begin = (warp >> 2) * NCC;
        uint8_t* sbuf = smem + Cfg::CST_OFF + warp * 2048;
        const bool has_r2 = p.resid2 != nullptr;
        if constexpr (Cfg::A2T) {
            // residual tiles fetched by TMA one chunk ahead into this warp's two 2 KB buffers (chunk sequence q = NCC * it + cc)
            uint8_t* rb1 = smem + Cfg::A2_OFF + warp * 4096;
            uint8_t* rb2 = rb1 + 2048;
            auto fetch_resid = [&](int tile, int c) {       // lane 0 only
                mbar_expect_tx(&rbar[warp], has_r2 ? 4096 : 2048);
                tma_load_2d(rb1, &tmX, &rbar[warp], c * 16, tile * FW_BM + quad * 32);
                if (has_r2) tma_load_2d(rb2, &tmR2, &rbar[warp], c * 16, tile * FW_BM + quad * 32);
            };
            if (lane == 0) fetch_resid(blockIdx.x, c_begin);
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                if (warp == 0 && lane == 0) FW_TRACE(0, it, 0);
                mbar_wait_parked(&y_full[0], (uint32_t)(it & 1));
                if (warp == 0 && lane == 0) FW_TRACE(0, it, 1);
                tc_fence_after();
#pragma unroll 1
                for (int cc = 0; cc < NCC; ++cc) {           // 16-column chunks
                    const int c = c_begin + cc;
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(tmem_base + c * 16 + ((uint32_t)(quad * 32) << 16), v);
                    mbar_wait(&rbar[warp], (uint32_t)((it * NCC + cc) & 1));
                    const int sw = (lane >> 1) & 3;          // SWIZZLE_64B: 16-byte unit index ^= (row >> 1) & 3
                    float4 r1[4], r2[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        r1[j] = *reinterpret_cast<const float4*>(rb1 + lane * 64 + ((j ^ sw) << 4));
                        r2[j] = has_r2 ? *reinterpret_cast<const float4*>(rb2 + lane * 64 + ((j ^ sw) << 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    fence_proxy_async_smem();                // order this lane's generic-proxy reads before the async-proxy (TMA) overwrite
                    __syncwarp();                            // every lane has read the buffers: the next fetch may overwrite them
                    if (lane == 0) {
                        if (cc + 1 < NCC) fetch_resid(tile, c + 1);
                        else if (tile + (int)gridDim.x < num_tiles) fetch_resid(tile + gridDim.x, c_begin);
                    }
                    tmem_ld_wait();
                    if (cc == NCC - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&y_free[0]);
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
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
                    uint8_t* rowp = sbuf + lane * 64;
#pragma unroll
                    for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(rowp + ((j ^ sw) << 4)) = o[j];
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmOut, sbuf, c * 16, tile * FW_BM + quad * 32);
                        tma_store_commit();
                    }
                }
                if (warp == 0 && lane == 0) FW_TRACE(0, it, 2);
            }
        } else {
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int yb = it % NYB;
            const uint32_t ypar = (it / NYB) & 1;
            const long long row = (long long)tile * FW_BM + quad * 32 + lane;
            const bool row_ok = row < p.M;
            float4 r1[4], r2[4];
            auto load_resid = [&](int c) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { r1[j] = make_float4(0.f, 0.f, 0.f, 0.f); r2[j] = r1[j]; }
                if (row_ok) {
                    const float4* rp = reinterpret_cast<const float4*>(p.x + row * C + c * 16);
#pragma unroll
                    for (int j = 0; j < 4; ++j) r1[j] = rp[j];
                    if (has_r2) {
                        const float4* rq = reinterpret_cast<const float4*>(p.resid2 + row * C + c * 16);
#pragma unroll
                        for (int j = 0; j < 4; ++j) r2[j] = rq[j];
                    }
                }
            };
            load_resid(c_begin);                             // in flight while fc2 of this tile finishes
            if (warp == 0 && lane == 0) FW_TRACE(0, it, 0);
            mbar_wait_parked(&y_full[yb], ypar);
            if (warp == 0 && lane == 0) FW_TRACE(0, it, 1);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < NCC; ++cc) {               // 16-column chunks
                const int c = c_begin + cc;
                uint32_t v[16];
                tmem_ld_32x32b_x16(tmem_base + yb * 192 + c * 16 + ((uint32_t)(quad * 32) << 16), v);
                tmem_ld_wait();
                if (cc == NCC - 1) {
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
                if (cc + 1 < NCC) load_resid(c + 1);         // next chunk's residuals fly during the staging / TMA store below
                if (lane == 0) tma_store_wait_read<0>();     // the previous chunk's store has read the staging buffer
                __syncwarp();
                uint8_t* rowp = sbuf + lane * 64;
                const int sw = (lane >> 1) & 3;              // SWIZZLE_64B: 16-byte unit index ^= (row >> 1) & 3
#pragma unroll
                for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(rowp + ((j ^ sw) << 4)) = o[j];
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmOut, sbuf, c * 16, tile * FW_BM + quad * 32);
                    tma_store_commit();
                }
            }
            if (warp == 0 && lane == 0) FW_TRACE(0, it, 2);
        }
        }
        if (lane == 0) tma_store_wait_all<0>();
    } else if (warp == FW_W_TMA) {
        // ============================================================ weight ring producer
        long long n = 0;                                     // slot sequence number of this CTA
        for (int t = 0; t < my_tiles; ++t) {
#pragma unroll 1
            for (int e = 0; e < 2 * NCH; ++e) {
                const bool is_w2 = e >= 2 && (e == 2 * NCH - 1 || (e & 1) == 0);
                const int j = e < 2 ? e : (e == 2 * NCH - 1 ? NCH - 1 : (is_w2 ? (e - 2) >> 1 : (e + 1) >> 1));
#pragma unroll 1
                for (int s = 0; s < NS; ++s, ++n) {
                    const int slot = (int)(n % NSLOT);
                    const uint32_t par = (uint32_t)((n / NSLOT) & 1);
                    if (lane == 0) FW_TRACE(6, (int)n, 0);
                    mbar_wait_parked(&ring_empty[slot], par ^ 1);
                    if (lane == 0) FW_TRACE(6, (int)n, 1);
                    if (elect_one_sync()) {
                        uint8_t* dst = smem + Cfg::RING_OFF + slot * Cfg::SLOT;
                        mbar_expect_tx(&ring_full[slot], Cfg::SLOT);
                        if (is_w2) {
                            tma_load_2d(dst, &tmW2, &ring_full[slot], j * 64, s * Cfg::YW);          // W2[YW s.., 64 j..]
                        } else {
#pragma unroll
                            for (int kb = 0; kb < Cfg::KBS; ++kb)                                    // W1[64 j.., YW s + 64 kb..]
                                tma_load_2d(dst + kb * 8192, &tmW1, &ring_full[slot], s * Cfg::YW + kb * 64, j * 64);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == FW_W_MMA) {
        // ============================================================ fc1 MMA issue
        constexpr uint32_t idesc1 = umma_idesc_bf16(FW_BM, 64);
        const uint64_t dA1 = umma_desc_sw128(smem_u32(smem + Cfg::A1_OFF));
        const uint32_t ring_u32 = smem_u32(smem + Cfg::RING_OFF);
        long long g = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int ab = t % NA1;
            mbar_wait_parked(&a1_full[ab], (uint32_t)((t / NA1) & 1));   // this tile's LayerNorm output is in A1[ab]
#pragma unroll 1
            for (int j = 0; j < NCH; ++j, ++g) {
                const int hb = j & 1;
                if (lane == 0) FW_TRACE(1, (int)g, 0);
                mbar_wait_parked(&h_free[hb], (uint32_t)(((g >> 1) & 1) ^ 1));
                if (lane == 0) FW_TRACE(1, (int)g, 1);
                tc_fence_after();
                const int e = j < 2 ? j : 2 * j - 1;
                const long long n0 = ((long long)t * 2 * NCH + e) * NS;
#pragma unroll 1
                for (int s = 0; s < NS; ++s) {
                    const long long n = n0 + s;
                    const int slot = (int)(n % NSLOT);
                    mbar_wait_parked(&ring_full[slot], (uint32_t)((n / NSLOT) & 1));
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t d = tmem_base + FW_TM_H + hb * 64;
                        const uint64_t db = umma_desc_sw128(ring_u32 + slot * Cfg::SLOT);
#pragma unroll
                        for (int kb = 0; kb < Cfg::KBS; ++kb)   // descriptor start-address field is in 16-byte units
                            umma_f16_ss_run<4>(d, dA1 + (uint64_t)((ab * Cfg::KB1 + s * Cfg::KBS + kb) * (Cfg::A1_KB >> 4)), db + (uint64_t)(kb * (8192 >> 4)), idesc1,
                                               (s | kb) != 0);
                        umma_commit(&ring_empty[slot]);
                        if (s == NS - 1) {
                            umma_commit(&h_full[hb]);
                            if (j == NCH - 1) umma_commit(&a1_free[ab]);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) FW_TRACE(1, (int)g, 2 + s);
                }
            }
        }
    } else if (warp == FW_W_MMA2) {
        // ============================================================ fc2 MMA issue: Y[:, 192 s..] += A2_j W2[192 s.., 64 j..]^T
        constexpr uint32_t idesc2 = umma_idesc_f16(FW_BM, Cfg::YW);   // A2 (GELU output) and W2 are fp16
        const uint64_t dA2 = umma_desc_sw128(smem_u32(smem + Cfg::A2_OFF));
        const uint32_t ring_u32 = smem_u32(smem + Cfg::RING_OFF);
        long long g = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int yb = t % NYB;
#pragma unroll 1
            for (int j = 0; j < NCH; ++j, ++g) {
                const int b = j & 1;
                if (lane == 0) FW_TRACE(2, (int)g, 0);
                mbar_wait_parked(&a2_full[b], (uint32_t)((g >> 1) & 1));
                if (lane == 0) FW_TRACE(2, (int)g, 1);
                if (j == 0) mbar_wait_parked(&y_free[yb], (uint32_t)(((t / NYB) & 1) ^ 1));
                if (lane == 0) FW_TRACE(2, (int)g, 2);
                tc_fence_after();
                const int e = j < NCH - 1 ? 2 * j + 2 : 2 * NCH - 1;
                const long long n0 = ((long long)t * 2 * NCH + e) * NS;
#pragma unroll 1
                for (int s = 0; s < NS; ++s) {
                    const long long n = n0 + s;
                    const int slot = (int)(n % NSLOT);
                    mbar_wait_parked(&ring_full[slot], (uint32_t)((n / NSLOT) & 1));
                    tc_fence_after();
                    if (elect_one_sync()) {
                        if constexpr (Cfg::A2T)
                            umma_f16_ts_run4(tmem_base + yb * 192 + s * Cfg::YW, tmem_base + Cfg::TM_A2 + b * 32,
                                             umma_desc_sw128(ring_u32 + slot * Cfg::SLOT), idesc2, j != 0);
                        else
                            umma_f16_ss_run<4>(tmem_base + yb * 192 + s * Cfg::YW, dA2 + (uint64_t)(b * (Cfg::A2_BYTES >> 4)),
                                               umma_desc_sw128(ring_u32 + slot * Cfg::SLOT), idesc2, j != 0);
                        umma_commit(&ring_empty[slot]);
                        if (s == NS - 1) {
                            umma_commit(&a2_free[b]);
                            if (j == NCH - 1) umma_commit(&y_full[yb]);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) FW_TRACE(2, (int)g, 3 + s);
                }
            }
        }
    } else {
        // ============================================================ GELU + LayerNorm warps
        const int ew = warp - FW_W_GELU;
        const int quad = warp & 3;
        const int grp = ew >> 3;          // chunk parity this warp works on == H accumulator / A2 buffer it uses
        const int half = (ew >> 2) & 1;   // which 32 of the chunk's 64 hidden columns
        const int row = quad * 32 + lane; // row inside the tile == TMEM lane
        // LayerNorm mapping: 16 lanes per row (C/16 contiguous channels each), 2 rows per warp instruction, 8 rows per warp
        const int l16 = lane & 15, rsub = lane >> 4;
        uint8_t* a1 = smem + Cfg::A1_OFF;
        auto prefetch_tile = [&](int tile) {   // pull this warp's 8 rows of x (and of the second residual) into L2
            if (tile >= num_tiles) return;
            const long long r0 = (long long)tile * FW_BM + ew * 8;
            for (int ofs = lane * 128; ofs < 8 * C * 4; ofs += 32 * 128) {
                if ((r0 * C) * 4 + ofs + 128 > (long long)p.M * C * 4) break;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.x + r0 * C) + ofs));
                if (p.resid2 != nullptr)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.resid2 + r0 * C) + ofs));
            }
        };
        // LayerNorm of rows [ew*8, ew*8+8) of `tile` into A1 (bf16, K-major SWIZZLE_128B k-blocks); lt = CTA-local tile counter
        auto layernorm_tile = [&](int tile, int lt) {
            prefetch_tile(tile + gridDim.x);
            constexpr int V4 = Cfg::LN_V4, STEPS = Cfg::LN_STEPS;
            bool waited = false;
#pragma unroll 1
            for (int sp = 0; sp < 4 / STEPS; ++sp) {
                float4 v[STEPS][V4];
#pragma unroll
                for (int st = 0; st < STEPS; ++st) {
                    const long long r = (long long)tile * FW_BM + ew * 8 + (sp * STEPS + st) * 2 + rsub;
                    if (r < p.M) {
                        const float4* xr = reinterpret_cast<const float4*>(p.x + r * C + l16 * Cfg::LN_CH);
#pragma unroll
                        for (int q = 0; q < V4; ++q) v[st][q] = __ldg(xr + q);
                    } else {
#pragma unroll
                        for (int q = 0; q < V4; ++q) v[st][q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                if (!waited) {
                    if (ew == 0 && lane == 0) FW_TRACE(3, lt, 0);
                    mbar_wait_parked(&a1_free[lt % NA1], (uint32_t)(((lt / NA1) & 1) ^ 1));   // the fc1 MMAs that read this buffer have retired
                    if (ew == 0 && lane == 0) FW_TRACE(3, lt, 1);
                    waited = true;
                }
                float sm[STEPS], qv[STEPS];
#pragma unroll
                for (int st = 0; st < STEPS; ++st) {
                    sm[st] = 0.f;
#pragma unroll
                    for (int q = 0; q < V4; ++q) sm[st] += (v[st][q].x + v[st][q].y) + (v[st][q].z + v[st][q].w);
                }
#pragma unroll
                for (int sh = 8; sh > 0; sh >>= 1)
#pragma unroll
                    for (int st = 0; st < STEPS; ++st) sm[st] += __shfl_xor_sync(0xffffffffu, sm[st], sh);
#pragma unroll
                for (int st = 0; st < STEPS; ++st) {
                    const float mean = sm[st] * (1.0f / C);
                    qv[st] = 0.f;
#pragma unroll
                    for (int q = 0; q < V4; ++q) {
                        v[st][q].x -= mean; v[st][q].y -= mean; v[st][q].z -= mean; v[st][q].w -= mean;
                        qv[st] += (v[st][q].x * v[st][q].x + v[st][q].y * v[st][q].y) + (v[st][q].z * v[st][q].z + v[st][q].w * v[st][q].w);
                    }
                }
#pragma unroll
                for (int sh = 8; sh > 0; sh >>= 1)
#pragma unroll
                    for (int st = 0; st < STEPS; ++st) qv[st] += __shfl_xor_sync(0xffffffffu, qv[st], sh);
#pragma unroll
                for (int st = 0; st < STEPS; ++st) {
                    const float rstd = rsqrtf(qv[st] * (1.0f / C) + 1e-5f);
                    const int rr = ew * 8 + (sp * STEPS + st) * 2 + rsub;     // row inside the tile
#pragma unroll
                    for (int q = 0; q < V4; ++q) {
                        const int c0 = l16 * Cfg::LN_CH + q * 4;              // 4 channels = 8 bytes of bf16, inside one 16-byte unit
                        const float4 gm = *reinterpret_cast<const float4*>(gs + c0);
                        const float4 bt = *reinterpret_cast<const float4*>(bs + c0);
                        uint2 pk;
                        pk.x = pack_bf16x2(fmaf(v[st][q].x * rstd, gm.x, bt.x), fmaf(v[st][q].y * rstd, gm.y, bt.y));
                        pk.y = pack_bf16x2(fmaf(v[st][q].z * rstd, gm.z, bt.z), fmaf(v[st][q].w * rstd, gm.w, bt.w));
                        const int kb = c0 >> 6, cc = c0 & 63;
                        *reinterpret_cast<uint2*>(a1 + ((lt % NA1) * Cfg::KB1 + kb) * Cfg::A1_KB + rr * 128 + (((cc >> 3) ^ (rr & 7)) << 4) + (cc & 7) * 2) = pk;
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a1_full[lt % NA1]);
            if (ew == 0 && lane == 0) FW_TRACE(3, lt, 2);
        };

        prefetch_tile(blockIdx.x);
        if (my_tiles > 0) layernorm_tile(blockIdx.x, 0);
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
#pragma unroll 1
            for (int jj = 0; jj < NCH / 2; ++jj) {
                const int j = 2 * jj + grp;
                const long long g = (long long)it * NCH + j;             // global chunk index of this CTA
                if ((ew & 7) == 0 && lane == 0) FW_TRACE(4 + grp, (int)g, 0);
                mbar_wait_parked(&h_full[grp], (uint32_t)((g >> 1) & 1));
                if ((ew & 7) == 0 && lane == 0) FW_TRACE(4 + grp, (int)g, 1);
                tc_fence_after();
                uint32_t pk[16];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(tmem_base + FW_TM_H + grp * 64 + half * 32 + s * 16 + ((uint32_t)(quad * 32) << 16), v);
                    const float* bb = p.b1h + j * 64 + half * 32 + s * 16;   // uniform across the warp: L1 broadcast loads
                    float4 b4[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) b4[i] = __ldg(reinterpret_cast<const float4*>(bb) + i);
                    tmem_ld_wait();
                    if (s == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&h_free[grp]);
                    }
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const float4 hb4 = b4[i >> 2];
                        pk[s * 8 + i / 2] = gelu_erf_f16x2_halved(fmaf(__uint_as_float(v[i]), 0.5f, hb4.x), fmaf(__uint_as_float(v[i + 1]), 0.5f, hb4.y));
                        pk[s * 8 + i / 2 + 1] = gelu_erf_f16x2_halved(fmaf(__uint_as_float(v[i + 2]), 0.5f, hb4.z), fmaf(__uint_as_float(v[i + 3]), 0.5f, hb4.w));
                    }
                }
                if ((ew & 7) == 0 && lane == 0) FW_TRACE(4 + grp, (int)g, 2);
                mbar_wait_parked(&a2_free[grp], (uint32_t)(((g >> 1) & 1) ^ 1));   // fc2 MMAs that read the previous contents have retired
                tc_fence_after();                               // order the tcgen05.st below after those MMAs' reads of the A2 columns
                if ((ew & 7) == 0 && lane == 0) FW_TRACE(4 + grp, (int)g, 3);
                if constexpr (Cfg::A2T) {
                    // 32 fp16 of this lane's row = 16 packed columns of the A2 tile in tensor memory (K pair 2c, 2c+1 in column c)
                    tmem_st_32x32b_x16(tmem_base + Cfg::TM_A2 + grp * 32 + half * 16 + ((uint32_t)(quad * 32) << 16), pk);
                    tmem_st_wait();
                    tc_fence_before();
                } else {
                    uint8_t* rowp = smem + Cfg::A2_OFF + grp * Cfg::A2_BYTES + row * 128;
                    const int sw = row & 7;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<uint4*>(rowp + (((half * 4 + q) ^ sw) << 4)) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
                    fence_proxy_async_smem();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&a2_full[grp]);
                if ((ew & 7) == 0 && lane == 0) FW_TRACE(4 + grp, (int)g, 4);
                // two A1 buffers: the next tile's LayerNorm runs early in this tile, group 0 after its first chunk and group 1
                // after its second, in the shadow of the other group's GELU
                if (NA1 == 2 && jj == grp && tile + (int)gridDim.x < num_tiles) layernorm_tile(tile + gridDim.x, it + 1);
            }
            // one A1 buffer: the next tile's LayerNorm waits for the last fc1 MMA of this tile, which was issued when this
            // group's second-to-last chunk left TMEM. Group 0 gets here one chunk before group 1.
            if (NA1 == 1 && tile + (int)gridDim.x < num_tiles) layernorm_tile(tile + gridDim.x, it + 1);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == FW_W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int C>
static int launch_ffn_wide(const float* x, const float* resid2, float* out, long long M, const float* gamma, const float* beta,
                           const __nv_bfloat16* w1, const float* b1_half, const __half* w2_f16, const float* b2, int num_sms, cudaStream_t stream) {
    using Cfg = FwCfg<C>;
    CUtensorMap t1, t2, to;
    ARD_TRY(make_tmap_2d(&t1, w1, 2, C, Cfg::HD, (uint64_t)C * 2, 64, 64, 128));
    ARD_TRY(make_tmap_2d(&t2, w2_f16, 2, Cfg::HD, C, (uint64_t)Cfg::HD * 2, 64, Cfg::YW, 128));
    ARD_TRY(make_tmap_2d(&to, out, 4, C, (uint64_t)M, (uint64_t)C * 4, 16, 32, 64));
    CUtensorMap tx, tr;
    ARD_TRY(make_tmap_2d(&tx, x, 4, C, (uint64_t)M, (uint64_t)C * 4, 16, 32, 64));
    ARD_TRY(make_tmap_2d(&tr, resid2 ? resid2 : x, 4, C, (uint64_t)M, (uint64_t)C * 4, 16, 32, 64));
    static bool attr_set = false;
    if (!attr_set) {
        ARD_CUDA(cudaFuncSetAttribute(ffn_wide_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    FwParams p;
    p.x = x; p.resid2 = resid2; p.gamma = gamma; p.beta = beta; p.b1h = b1_half; p.b2 = b2; p.M = (int)M;
    const int tiles = (int)((M + FW_BM - 1) / FW_BM);
    const int grid = tiles < num_sms ? tiles : num_sms;
    const double MC = (double)M * C;
    ProfScope ps(PROF_FFN, stream, 2.0 * M * C * Cfg::HD * 2.0, MC * 4.0 * (2.0 + (resid2 ? 1.0 : 0.0)) + 2.0 * 2.0 * C * Cfg::HD);
    ARD_CUDA(enqueue_pdl(ffn_wide_kernel<C>, dim3(grid), dim3(FW_THREADS), Cfg::SMEM_BYTES, stream, t1, t2, to, tx, tr, p));
    return check_cuda(cudaGetLastError(), "ffn_wide launch");
}

// x_out = x + fc2(gelu(fc1(LN(x)))) (+ resid2), C = 192 or 384. x_out may alias x (each element is read by the warp that later
// writes it: the LayerNorm read of a tile precedes its output store by construction). b1_half = 0.5 * fc1 bias.
int ffn_fused_wide(const float* x, const float* resid2, float* out, long long M, int C, const float* gamma, const float* beta,
                   const __nv_bfloat16* w1, const float* b1_half, const __half* w2_f16, const float* b2, int num_sms, cudaStream_t stream) {
    if (M <= 0) return 0;
    if (M > 0x7fffffffLL) return set_error(ARD_ERR_SHAPE, "ffn_wide: too many rows");
    if (C == 128) return launch_ffn_wide<128>(x, resid2, out, M, gamma, beta, w1, b1_half, w2_f16, b2, num_sms, stream);
    if (C == 192) return launch_ffn_wide<192>(x, resid2, out, M, gamma, beta, w1, b1_half, w2_f16, b2, num_sms, stream);
    if (C == 256) return launch_ffn_wide<256>(x, resid2, out, M, gamma, beta, w1, b1_half, w2_f16, b2, num_sms, stream);
    if (C == 384) return launch_ffn_wide<384>(x, resid2, out, M, gamma, beta, w1, b1_half, w2_f16, b2, num_sms, stream);
    return set_error(ARD_ERR_SHAPE, "ffn_wide: C = %d not supported (128, 192, 256, 384)", C);
}

}  // namespace ard

#ifdef ARD_FFN_TRACE
extern "C" int ard_debug_ffw_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, ard::g_ffw_trace, sizeof(long long) * 8 * 64 * 8);
}
#endif
