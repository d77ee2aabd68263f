// Window-resident attention block of the 96-channel stage, one kernel:
//
//     y = x + proj'( window_attention( LayerNorm1(x) ) )        proj' = out-projection with the ResiDual fold (M W_p, (b_p - mu) M)
//
// Reference: SwinTransformerBlock.forward htsat.py:449-476 (norm1, roll, window_partition, WindowAttention :326-357,
// window_reverse, roll back, shortcut add) and the patched form src/residual.py:58-92 (ResiDual on the attention output,
// folded into the projection for the current lambda: ard_api.cu::ensure_fold).
//
// Unfused this is four kernels (ln_qkv, window_attention, proj GEMM) that move 30 B per token-channel through HBM: the bf16
// qkv tensor is written and re-read, so is the attention output. Here a CTA owns a PAIR of 8x8 windows (128 tokens = one
// UMMA M tile); x is read once, y written once (8 B per token-channel) and every contraction runs on tcgen05 with its
// accumulator in tensor memory:
//
//   A   = LayerNorm1(x rows of the two windows, gathered with the cyclic shift)      bf16, SWIZZLE_64B K-major, smem
//   QK  = A  Wqk_pad^T      M128 N256 K96   per-head padded layout: head h -> columns 32h..32h+23 (hd 24 -> 32, zero rows),
//                                           q rows carry head_dim^-0.5 * log2(e)
//   V^T = Wv_pad A^T        M128 N128 K96   computed TRANSPOSED (weights as the A operand, tokens as N) so that the epilogue
//                                           thread of channel c writes row c of V^T[channel, key]: the K-major B operand of P V
//   S_h = Q_h K_h^T         M128 N128 K32   both windows at once; the two cross-window 64x64 blocks are computed and ignored
//   P_h = softmax(S_h + rel-pos bias + shift mask)   one thread per query row (TMEM lane): 64 logits in registers, exp2,
//                                           probabilities written back to TENSOR MEMORY as the bf16 A operand (zeros in the
//                                           cross-window half), S double-buffered so head h+1's S MMA overlaps head h's softmax
//   O_h = P_h V_h           M128 N32  K128  A from TMEM, B = rows 32h.. of V^T
//   Y   = O_pad Wp_pad^T    M128 N96  K128  + b' + x -> y (fp32, token order: window_reverse / roll back are address arithmetic)
//
// TMEM (512 columns): QK accumulator 0-255, V^T accumulator 256-383, O 384-511; after the QKV drain S[0], S[1] reuse 0-255,
// P[0], P[1] reuse 256-383 (64 columns each: 128 bf16 per row); Y reuses the O columns (384-479) once O is in shared memory, so
// the NEXT tile's QK / V^T MMAs (columns 0-383) run underneath this tile's output epilogue.
// Row sums ride on the tensor core: padded channel 24 of every head of V^T is set to ones, so column 24 of O_h is sum_j P~_ij of
// the bf16 probabilities actually multiplied; the O drain divides by it (no per-element sum / normalise in the softmax).
// Warp roles (13 warps): 0-7 tensor-memory warps (lane quadrant w & 3, half w >> 2: drain QKV, softmax of heads {half, half+2},
// drain O, output epilogue), 8 = TMA weights + MMA issue (warp-convergent, elected lane), 9-12 = LayerNorm of the NEXT tile
// (it overlaps the attention phase: the A tile is free as soon as the QK / V^T MMAs have read it).
#include "ard_common.cuh"
#include "ard_handle.h"

namespace ard {

// Development aid (tools/ab_trace.py builds a separate library with -DARD_AB_TRACE): per-role clock64() stamps of CTA 0.
#ifdef ARD_AB_TRACE
__device__ long long g_ab_trace[4][32][24];   // [role][tile index][event]
#define AB_TRACE(role, idx, field) do { if (blockIdx.x == 0 && lane == 0 && (idx) < 32) g_ab_trace[role][idx][field] = clock64(); } while (0)
#else
#define AB_TRACE(role, idx, field) do { } while (0)
#endif

constexpr int AB_C = 96, AB_NH = 4, AB_HD = 24;
constexpr int AB_TM_WARPS = 8, AB_W_MMA = 8, AB_W_LN = 9, AB_LN_WARPS = 4;
constexpr int AB_THREADS = (AB_W_LN + AB_LN_WARPS) * 32;   // 416

constexpr int AB_KB = 128 * 64;                    // bytes of a 128-row x 32-element (64 B) SWIZZLE_64B k-block
constexpr int AB_WQK_OFF = 0, AB_WQK_KB = 256 * 64;               // 3 k-blocks of [256 rows x 64 B]
constexpr int AB_WV_OFF = AB_WQK_OFF + 3 * AB_WQK_KB;             // 49152: 3 k-blocks of [128 x 64 B]
constexpr int AB_WP_OFF = AB_WV_OFF + 3 * AB_KB, AB_WP_KB = 96 * 64;   // 73728: 4 k-blocks of [96 x 64 B]
constexpr int AB_A_OFF = AB_WP_OFF + 4 * AB_WP_KB;                // 98304: 3 k-blocks
constexpr int AB_Q_OFF = AB_A_OFF + 3 * AB_KB;                    // 122880: 4 k-blocks (one per head); later the O_pad operand
constexpr int AB_K_OFF = AB_Q_OFF + 4 * AB_KB;                    // 155648
constexpr int AB_VT_OFF = AB_K_OFF + 4 * AB_KB;                   // 188416: 4 k-blocks of 32 keys, rows = padded channel
constexpr int AB_VEC_OFF = AB_VT_OFF + 4 * AB_KB;                 // 221184: bq[128] bp[96] gamma[96] beta[96]
constexpr int AB_TAB_STRIDE = 24, AB_TAB_N = 15 * AB_TAB_STRIDE;  // rel-pos bias table per head as [15][24]: index (dy+7)*24 + (dx+7); the
                                                                  // stride 24 puts the 32 rows of a warp (4 ty x 8 tx) on 32 different banks
constexpr int AB_TAB_OFF = AB_VEC_OFF + (128 + 3 * 96) * 4;       // [4][360] floats (x log2 e)
constexpr int AB_ROW_OFF = AB_TAB_OFF + 4 * AB_TAB_N * 4;         // int[128]: row of x / out of each token of the tile
constexpr int AB_BAR_OFF = AB_ROW_OFF + 128 * 4;
constexpr int AB_SMEM_BYTES = AB_BAR_OFF + 256 + 1024;
constexpr int AB_STAGE_OFF = AB_Q_OFF, AB_STAGE_LD = 400;         // output staging [128][100] fp32 over the (by then dead) Q / K tiles
static_assert(128 * AB_STAGE_LD <= 8 * AB_KB, "attn_block: staging tile must fit in the Q + K tiles");
static_assert(AB_SMEM_BYTES <= 227 * 1024, "attn_block: shared memory budget");

constexpr uint32_t AB_TM_QK = 0, AB_TM_VT = 256, AB_TM_O = 384, AB_TM_S = 0, AB_TM_P = 256, AB_TM_Y = 384;
constexpr float AB_LOG2E = 1.4426950408889634f;

struct AttnBlockParams {
    const float* x;       // [B*R*R, 96] fp32, token order
    float* out;           // [B*R*R, 96] fp32
    const float* bq;      // [128] padded q bias (scaled). The k bias cancels in the softmax (constant per query row) and the v bias
                          // passes through P (rows sum to 1): it is folded into bp on the host side (attn_block_pad_proj)
    const float* bp;      // [96]  (folded) projection bias + Wp' bv
    const float* gamma;   // norm1
    const float* beta;
    const float* table;   // [4][15][24] relative-position bias x log2(e)
    int R, shift, n_tiles;   // tokens per side (64), cyclic shift (0 / 4), number of window pairs
};

// D (+)= A_tmem * B for ONE k-step (A: 128 lanes x 8 columns of packed bf16 pairs)
ARD_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
ARD_DEVINL float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// byte offset of element (row r, column c) in a SWIZZLE_64B K-major tile made of k-blocks of `kb_bytes`
ARD_DEVINL int sw64_off(int r, int c, int kb_bytes) {
    const int kb = c >> 5, cc = c & 31;
    return kb * kb_bytes + r * 64 + ((((cc >> 3) ^ (r >> 1)) & 3) << 4) + (cc & 7) * 2;
}

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_block_kernel(const __grid_constant__ CUtensorMap tmWqk, const __grid_constant__ CUtensorMap tmWv, const __grid_constant__ CUtensorMap tmWp,
                  const AttnBlockParams p) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* bq_s = reinterpret_cast<float*>(smem + AB_VEC_OFF);
    float* bp_s = bq_s + 128;
    float* g_s = bp_s + 96;
    float* b_s = g_s + 96;
    float* tab_s = reinterpret_cast<float*>(smem + AB_TAB_OFF);
    int* rowidx_s = reinterpret_cast<int*>(smem + AB_ROW_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AB_BAR_OFF);
    uint64_t* w_full = bars + 0;
    uint64_t* a_full = bars + 1;
    uint64_t* a_free = bars + 2;
    uint64_t* acc_full = bars + 3;
    uint64_t* qkv_ready = bars + 4;
    uint64_t* s_full = bars + 5;    // [2]
    uint64_t* s_free = bars + 7;    // [2]
    uint64_t* p_full = bars + 9;    // [2]
    uint64_t* p_free = bars + 11;   // [2]
    uint64_t* o_full = bars + 13;
    uint64_t* ao_ready = bars + 14;
    uint64_t* y_full = bars + 15;
    uint64_t* y_free = bars + 16;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int R = p.R, nWr = R >> 3, nW = nWr * nWr;

    pdl_wait();   // the folded projection bias / weights may have been written by the kernel just before this one (lambda update)
    for (int i = threadIdx.x; i < 128; i += AB_THREADS) bq_s[i] = p.bq[i];
    for (int i = threadIdx.x; i < 96; i += AB_THREADS) {
        bp_s[i] = p.bp[i];
        g_s[i] = p.gamma[i];
        b_s[i] = p.beta[i];
    }
    for (int i = threadIdx.x; i < 4 * AB_TAB_N; i += AB_THREADS) tab_s[i] = p.table[i];
    if (warp == AB_W_MMA && lane == 0) {
        tma_prefetch_desc(&tmWqk);
        tma_prefetch_desc(&tmWv);
        tma_prefetch_desc(&tmWp);
        mbar_init(w_full, 1);
        mbar_init(a_full, AB_LN_WARPS);
        mbar_init(a_free, 1);
        mbar_init(acc_full, 1);
        mbar_init(qkv_ready, AB_TM_WARPS);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 4);
            mbar_init(&p_full[i], 4);
            mbar_init(&p_free[i], 1);
        }
        mbar_init(o_full, 1);
        mbar_init(ao_ready, AB_TM_WARPS);
        mbar_init(y_full, 1);
        mbar_init(y_free, AB_TM_WARPS);
        fence_barrier_init();
    }
    if (warp == AB_W_MMA) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // token (row r of tile `tile`) -> row index of x / out: window pair -> (clip, wy, wx), roll by -shift (htsat.py:452-460)
    auto token_row = [&](int tile, int r) -> long long {
        const int widx = tile * 2 + (r >> 6);
        const int b = widx / nW, win = widx - b * nW;
        const int wy = win / nWr, wx = win - wy * nWr;
        const int i = r & 63;
        int y = wy * 8 + (i >> 3) + p.shift, x = wx * 8 + (i & 7) + p.shift;
        if (y >= R) y -= R;
        if (x >= R) x -= R;
        return ((long long)b * R + y) * R + x;
    };

    if (warp < AB_TM_WARPS) {
        // ============================================================ tensor-memory warps
        const int quad = warp & 3, half = warp >> 2;
        const int row = quad * 32 + lane;                 // TMEM lane == row of the tile (token, or padded channel for V^T)
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int sw = (row >> 1) & 3;
        const int win_in_tile = quad >> 1;
        const int qi = row & 63, ty = qi >> 3, tx = qi & 7;
        const int ci = ty * AB_TAB_STRIDE + tx + 7 * AB_TAB_STRIDE + 7;   // bias index = ci - (jy * 24 + jx)   (relative_position_index, htsat.py:301-316)
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            // ---- 1. drain the QK / V^T accumulators into their shared-memory operand tiles (bf16)
            const int trole = (quad == 0) ? half : 3;     // trace: warps 0 and 4
            if (quad == 0) AB_TRACE(trole, it, 0);
            mbar_wait_parked(acc_full, (uint32_t)(it & 1));
            if (quad == 0) AB_TRACE(trole, it, 1);
            tc_fence_after();
            {
                auto put = [&](const uint32_t (&v)[32], uint8_t* rowp, const float* bb, bool ones) {   // 32 fp32 (+ bias) -> one 64-byte bf16 row, SWIZZLE_64B
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t pk[4];
                        float bq8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        if (bb != nullptr) {                  // two broadcast LDS.128 instead of eight LDS.32: the drain is shared-memory-pipe bound
                            const float4 b0 = *reinterpret_cast<const float4*>(bb + q * 8), b1 = *reinterpret_cast<const float4*>(bb + q * 8 + 4);
                            bq8[0] = b0.x; bq8[1] = b0.y; bq8[2] = b0.z; bq8[3] = b0.w; bq8[4] = b1.x; bq8[5] = b1.y; bq8[6] = b1.z; bq8[7] = b1.w;
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float a0 = __uint_as_float(v[q * 8 + 2 * k]) + bq8[2 * k], a1 = __uint_as_float(v[q * 8 + 2 * k + 1]) + bq8[2 * k + 1];
                            pk[k] = ones ? 0x3F803F80u : pack_bf16x2(a0, a1);
                        }
                        *reinterpret_cast<uint4*>(rowp + ((q ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                };
#pragma unroll
                for (int hq = 0; hq < 2; ++hq) {          // heads 2 half, 2 half + 1: q (with bias) and k chunks of this thread's row
                    const int h = half * 2 + hq;
                    uint32_t vq[32], vk[32];
                    tmem_ld_32x32b_x32(lane_addr + AB_TM_QK + h * 32, vq);
                    tmem_ld_32x32b_x32(lane_addr + AB_TM_QK + 128 + h * 32, vk);
                    tmem_ld_wait();
                    put(vq, smem + AB_Q_OFF + h * AB_KB + row * 64, bq_s + h * 32, false);
                    put(vk, smem + AB_K_OFF + h * AB_KB + row * 64, nullptr, false);
                }
                {                                         // V^T: lane = padded channel, columns = tokens (keys)
                    uint32_t v0[32], v1[32];
                    tmem_ld_32x32b_x32(lane_addr + AB_TM_VT + half * 64, v0);
                    tmem_ld_32x32b_x32(lane_addr + AB_TM_VT + half * 64 + 32, v1);
                    tmem_ld_wait();
                    const bool ones = (row & 31) == AB_HD;    // padded channel 24 of each head: the all-ones row that makes O_h[:, 24] the row sum
                    put(v0, smem + AB_VT_OFF + (half * 2) * AB_KB + row * 64, nullptr, ones);
                    put(v1, smem + AB_VT_OFF + (half * 2 + 1) * AB_KB + row * 64, nullptr, ones);
                }
            }
            if (half == 0) rowidx_s[row] = (int)token_row(tile, row);   // read by the output epilogue (after a bar.sync of these warps)
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(qkv_ready);
            if (quad == 0) AB_TRACE(trole, it, 2);

            // ---- 2. softmax of heads {half, half + 2}: one query row per thread
            const int widx = tile * 2 + win_in_tile;
            const int win = widx % nW;
            const bool rowmask = p.shift > 0 && (win / nWr) == nWr - 1;   // windows that straddle the roll seam (htsat.py:414-433)
            const bool colmask = p.shift > 0 && (win % nWr) == nWr - 1;
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
                const int h = 2 * hh + half;
                const uint32_t u = (uint32_t)(2 * it + hh);
                mbar_wait_parked(&s_full[half], u & 1);
                if (quad == 0) AB_TRACE(trole, it, 3 + 5 * hh);
                tc_fence_after();
                uint32_t v[64];
                {
                    uint32_t t0[32], t1[32];
                    tmem_ld_32x32b_x32(lane_addr + AB_TM_S + half * 128 + win_in_tile * 64, t0);
                    tmem_ld_32x32b_x32(lane_addr + AB_TM_S + half * 128 + win_in_tile * 64 + 32, t1);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) { v[j] = t0[j]; v[32 + j] = t1[j]; }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_free[half]);
                if (quad == 0) AB_TRACE(trole, it, 4 + 5 * hh);
                const float* tb = tab_s + h * AB_TAB_N + ci;
                float s[64];
                float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains: 16-deep instead of 64-deep
                if (rowmask || colmask) {
                    const bool my_y = ty >= 4, my_x = tx >= 4;
#pragma unroll
                    for (int j = 0; j < 64; ++j) {
                        const bool masked = (rowmask && (my_y != (j >= 32))) || (colmask && (my_x != ((j & 7) >= 4)));
                        s[j] = __uint_as_float(v[j]) + tb[-((j >> 3) * AB_TAB_STRIDE + (j & 7))] + (masked ? -100.0f * AB_LOG2E : 0.0f);
                        mx[j & 3] = fmaxf(mx[j & 3], s[j]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 64; ++j) {
                        s[j] = __uint_as_float(v[j]) + tb[-((j >> 3) * AB_TAB_STRIDE + (j & 7))];
                        mx[j & 3] = fmaxf(mx[j & 3], s[j]);
                    }
                }
                const float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
                uint32_t pk[32];                               // un-normalised probabilities (<= 1); the row sum comes back in O_h[:, 24]
#pragma unroll
                for (int j = 0; j < 32; ++j) pk[j] = pack_bf16x2(ex2_approx(s[2 * j] - m), ex2_approx(s[2 * j + 1] - m));
                if (quad == 0) AB_TRACE(trole, it, 5 + 5 * hh);
                mbar_wait_parked(&p_free[half], (u & 1) ^ 1);     // the P V MMAs that read the previous contents have retired
                if (quad == 0) AB_TRACE(trole, it, 6 + 5 * hh);
                tc_fence_after();
                const uint32_t pbase = lane_addr + AB_TM_P + half * 64;
                const uint32_t own = pbase + win_in_tile * 32, other = pbase + (win_in_tile ^ 1) * 32;
                {
                    uint32_t a[16], z[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) { a[j] = pk[j]; z[j] = 0u; }
                    tmem_st_32x32b_x16(own, a);
                    tmem_st_32x32b_x16(other, z);
                    tmem_st_32x32b_x16(other + 16, z);
#pragma unroll
                    for (int j = 0; j < 16; ++j) a[j] = pk[16 + j];
                    tmem_st_32x32b_x16(own + 16, a);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[half]);
                if (quad == 0) AB_TRACE(trole, it, 7 + 5 * hh);
            }

            // ---- 3. drain O (heads 2 half, 2 half + 1) -> O_pad operand of the projection (reuses the Q tile)
            mbar_wait_parked(o_full, (uint32_t)(it & 1));
            if (quad == 0) AB_TRACE(trole, it, 13);
            tc_fence_after();
#pragma unroll 1
            for (int hq = 0; hq < 2; ++hq) {
                const int h = half * 2 + hq;
                uint32_t v[32];
                tmem_ld_32x32b_x32(lane_addr + AB_TM_O + h * 32, v);
                tmem_ld_wait();
                const float inv = 1.0f / __uint_as_float(v[AB_HD]);   // sum_j P~_ij (>= the max term, 1)
                uint8_t* rowp = smem + AB_Q_OFF + h * AB_KB + row * 64;
#pragma unroll
                for (int q = 0; q < 3; ++q) {                 // 24 channels = three 16-byte units; the fourth (padding) is zero
                    uint32_t pk[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) pk[k] = pack_bf16x2(__uint_as_float(v[q * 8 + 2 * k]) * inv, __uint_as_float(v[q * 8 + 2 * k + 1]) * inv);
                    *reinterpret_cast<uint4*>(rowp + ((q ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
                *reinterpret_cast<uint4*>(rowp + ((3 ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ao_ready);
            if (quad == 0) AB_TRACE(trole, it, 14);

            // ---- 4. output: y = Y + b' + x. Each thread adds the bias to 48 channels of its row (columns [32 half, +32) and
            // [64 + 16 half, +16)) and parks them in a padded fp32 staging tile over the dead Q / K tiles; the 256 threads then
            // walk the tile in row-major float4 order so that x is read and y written in 384-byte coalesced runs (a row-per-thread
            // LDG / STG touches 32 different lines per instruction: measured 3.7 k cycles each per tile, 40 % of the first version).
            // The x loads are issued before the wait for the projection MMA, so their L2 latency hides behind it.
            asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight tensor-memory warps: row indices visible
            float4 xv[12];
            int goff[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const int f = (int)threadIdx.x + 256 * i;       // float4 index in the [128][24] tile
                const int r = f / 24, c4 = f - r * 24;
                goff[i] = rowidx_s[r] * AB_C + c4 * 4;
                xv[i] = __ldg(reinterpret_cast<const float4*>(p.x + goff[i]));   // shortcut (an L2 hit: read by the LayerNorm warps)
            }
            const int cA = 32 * half, cB = 64 + 16 * half;
            if (quad == 0) AB_TRACE(trole, it, 15);
            mbar_wait_parked(y_full, (uint32_t)(it & 1));
            if (quad == 0) AB_TRACE(trole, it, 16);
            tc_fence_after();
            {
                uint32_t ya[32], yb[16];
                tmem_ld_32x32b_x32(lane_addr + AB_TM_Y + cA, ya);
                tmem_ld_32x32b_x16(lane_addr + AB_TM_Y + cB, yb);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(y_free);
                uint8_t* srow = smem + AB_STAGE_OFF + row * AB_STAGE_LD;
#pragma unroll
                for (int j = 0; j < 12; ++j) {
                    const uint32_t* src = j < 8 ? &ya[j * 4] : &yb[(j - 8) * 4];
                    const int c = j < 8 ? cA + j * 4 : cB + (j - 8) * 4;
                    const float4 b4 = *reinterpret_cast<const float4*>(bp_s + c);
                    *reinterpret_cast<float4*>(srow + c * 4) = make_float4(__uint_as_float(src[0]) + b4.x, __uint_as_float(src[1]) + b4.y,
                                                                           __uint_as_float(src[2]) + b4.z, __uint_as_float(src[3]) + b4.w);
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");   // staging tile complete
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const int f = (int)threadIdx.x + 256 * i;
                const int r = f / 24, c4 = f - r * 24;
                const float4 yv = *reinterpret_cast<const float4*>(smem + AB_STAGE_OFF + r * AB_STAGE_LD + c4 * 16);
                *reinterpret_cast<float4*>(p.out + goff[i]) = make_float4(xv[i].x + yv.x, xv[i].y + yv.y, xv[i].z + yv.z, xv[i].w + yv.w);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");   // staging tile read out: the next tile's drain may overwrite Q / K
            if (quad == 0) AB_TRACE(trole, it, 17);
        }
    } else if (warp == AB_W_MMA) {
        // ============================================================ weight load + MMA issue (warp-convergent, elected lane)
        if (elect_one_sync()) {
            mbar_expect_tx(w_full, 3 * AB_WQK_KB + 3 * AB_KB + 4 * AB_WP_KB);
            for (int kb = 0; kb < 3; ++kb) {
                tma_load_2d(smem + AB_WQK_OFF + kb * AB_WQK_KB, &tmWqk, w_full, kb * 32, 0);
                tma_load_2d(smem + AB_WV_OFF + kb * AB_KB, &tmWv, w_full, kb * 32, 0);
            }
            for (int kb = 0; kb < 4; ++kb) tma_load_2d(smem + AB_WP_OFF + kb * AB_WP_KB, &tmWp, w_full, kb * 32, 0);
        }
        __syncwarp();
        mbar_wait(w_full, 0);
        constexpr uint32_t id_qk = umma_idesc_bf16(128, 256), id_vt = umma_idesc_bf16(128, 128), id_s = umma_idesc_bf16(128, 128),
                           id_pv = umma_idesc_bf16(128, 32), id_y = umma_idesc_bf16(128, 96);
        const uint64_t dWqk = umma_desc_sw64(smem_u32(smem + AB_WQK_OFF)), dWv = umma_desc_sw64(smem_u32(smem + AB_WV_OFF)),
                       dWp = umma_desc_sw64(smem_u32(smem + AB_WP_OFF)), dA = umma_desc_sw64(smem_u32(smem + AB_A_OFF)),
                       dQ = umma_desc_sw64(smem_u32(smem + AB_Q_OFF)), dK = umma_desc_sw64(smem_u32(smem + AB_K_OFF)),
                       dVT = umma_desc_sw64(smem_u32(smem + AB_VT_OFF));
        auto issue_qkv = [&](int it) {                       // QK and V^T accumulators of local tile `it` (TMEM columns 0-383)
            mbar_wait_parked(a_full, (uint32_t)(it & 1));
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
                for (int kb = 0; kb < 3; ++kb)
                    umma_f16_ss_run<2>(tmem_base + AB_TM_QK, dA + (uint64_t)(kb * (AB_KB >> 4)), dWqk + (uint64_t)(kb * (AB_WQK_KB >> 4)), id_qk, kb != 0);
#pragma unroll
                for (int kb = 0; kb < 3; ++kb)
                    umma_f16_ss_run<2>(tmem_base + AB_TM_VT, dWv + (uint64_t)(kb * (AB_KB >> 4)), dA + (uint64_t)(kb * (AB_KB >> 4)), id_vt, kb != 0);
                umma_commit(acc_full);
                umma_commit(a_free);
            }
            __syncwarp();
        };
        if (my_tiles > 0) issue_qkv(0);
        for (int it = 0; it < my_tiles; ++it) {
            AB_TRACE(2, it, 3);
            mbar_wait_parked(qkv_ready, (uint32_t)(it & 1));
            AB_TRACE(2, it, 4);
            tc_fence_after();
            if (elect_one_sync()) {
                for (int h = 0; h < 2; ++h) {
                    umma_f16_ss_run<2>(tmem_base + AB_TM_S + h * 128, dQ + (uint64_t)(h * (AB_KB >> 4)), dK + (uint64_t)(h * (AB_KB >> 4)), id_s, 0);
                    umma_commit(&s_full[h]);
                }
            }
            __syncwarp();
            mbar_wait_parked(y_free, (uint32_t)((it & 1) ^ 1));       // the previous tile's Y has left the O columns
#pragma unroll 1
            for (int h = 0; h < 4; ++h) {
                const int b = h & 1;
                const uint32_t u = (uint32_t)(2 * it + (h >> 1));
                if (h + 2 < 4) {                                     // S of head h+2 goes where head h's logits were: they are in registers by now
                    mbar_wait_parked(&s_free[b], u & 1);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        umma_f16_ss_run<2>(tmem_base + AB_TM_S + b * 128, dQ + (uint64_t)((h + 2) * (AB_KB >> 4)), dK + (uint64_t)((h + 2) * (AB_KB >> 4)), id_s, 0);
                        umma_commit(&s_full[b]);
                    }
                    __syncwarp();
                }
                AB_TRACE(2, it, 5 + 3 * h);
                mbar_wait_parked(&p_full[b], u & 1);
                AB_TRACE(2, it, 6 + 3 * h);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t d = tmem_base + AB_TM_O + h * 32, a = tmem_base + AB_TM_P + b * 64;
                    const uint64_t dv = dVT + (uint64_t)((h * 32 * 64) >> 4);   // rows 32h.. of every 32-key k-block of V^T
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_bf16_ts(d, a + ks * 8, dv + (uint64_t)((ks >> 1) * (AB_KB >> 4) + (ks & 1) * 2), id_pv, ks != 0);
                    umma_commit(&p_free[b]);
                    if (h == 3) umma_commit(o_full);
                }
                __syncwarp();
                AB_TRACE(2, it, 7 + 3 * h);
            }
            mbar_wait_parked(ao_ready, (uint32_t)(it & 1));
            AB_TRACE(2, it, 17);
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
                    umma_f16_ss_run<2>(tmem_base + AB_TM_Y, dQ + (uint64_t)(kb * (AB_KB >> 4)), dWp + (uint64_t)(kb * (AB_WP_KB >> 4)), id_y, kb != 0);
                umma_commit(y_full);
            }
            __syncwarp();
            AB_TRACE(2, it, 18);
            // the next tile's QK / V^T MMAs run underneath this tile's output epilogue: S / P (columns 0-383) are dead (every P V MMA
            // was issued after its p_full and executes before these), Y lives in the O columns
            if (it + 1 < my_tiles) issue_qkv(it + 1);
        }
    } else {
        // ============================================================ LayerNorm warps: 32 rows each, 8 lanes per row (12 channels per lane)
        const int lw = warp - AB_W_LN;
        const int l8 = lane & 7, rsub = lane >> 3;
        uint8_t* a1 = smem + AB_A_OFF;
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            if (it + 1 < my_tiles) {                      // pull the next tile's 32 rows of this warp into L2: 3 lines of 128 B per row
                const char* nr = reinterpret_cast<const char*>(p.x + token_row(tile + gridDim.x, lw * 32 + lane) * AB_C);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nr));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nr + 128));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nr + 256));
            }
#pragma unroll 1
            for (int bt = 0; bt < 2; ++bt) {              // two batches of 16 rows (4 row groups of 4)
                float4 v[4][3];
                int rr[4];
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                    rr[gi] = lw * 32 + bt * 16 + gi * 4 + rsub;
                    const float4* xr = reinterpret_cast<const float4*>(p.x + token_row(tile, rr[gi]) * AB_C + l8 * 12);
                    v[gi][0] = __ldg(xr); v[gi][1] = __ldg(xr + 1); v[gi][2] = __ldg(xr + 2);
                }
                if (lw == 0 && bt == 0) AB_TRACE(3, it, 20);
                if (bt == 0) mbar_wait_parked(a_free, (uint32_t)((it & 1) ^ 1));   // the previous tile's QK / V^T MMAs have read A
                if (lw == 0 && bt == 0) AB_TRACE(3, it, 21);
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                    float sm = 0.f;
#pragma unroll
                    for (int q = 0; q < 3; ++q) sm += (v[gi][q].x + v[gi][q].y) + (v[gi][q].z + v[gi][q].w);
#pragma unroll
                    for (int sh = 4; sh > 0; sh >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, sh);
                    const float mean = sm * (1.0f / AB_C);
                    float qv = 0.f;
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        v[gi][q].x -= mean; v[gi][q].y -= mean; v[gi][q].z -= mean; v[gi][q].w -= mean;
                        qv += (v[gi][q].x * v[gi][q].x + v[gi][q].y * v[gi][q].y) + (v[gi][q].z * v[gi][q].z + v[gi][q].w * v[gi][q].w);
                    }
#pragma unroll
                    for (int sh = 4; sh > 0; sh >>= 1) qv += __shfl_xor_sync(0xffffffffu, qv, sh);
                    const float rstd = rsqrtf(qv * (1.0f / AB_C) + 1e-5f);
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const int c0 = l8 * 12 + q * 4;
                        const float4 gm = *reinterpret_cast<const float4*>(g_s + c0);
                        const float4 bta = *reinterpret_cast<const float4*>(b_s + c0);
                        uint2 pk;
                        pk.x = pack_bf16x2(fmaf(v[gi][q].x * rstd, gm.x, bta.x), fmaf(v[gi][q].y * rstd, gm.y, bta.y));
                        pk.y = pack_bf16x2(fmaf(v[gi][q].z * rstd, gm.z, bta.z), fmaf(v[gi][q].w * rstd, gm.w, bta.w));
                        *reinterpret_cast<uint2*>(a1 + sw64_off(rr[gi], c0, AB_KB)) = pk;
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
            if (lw == 0) AB_TRACE(3, it, 22);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == AB_W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ host side
// Wp_pad[n][32 h + d] = Wp[n][24 h + d] (zero elsewhere): the projection consumes the per-head padded O layout.
// bp_eff[n] = bp[n] + sum_c Wp[n][c] bv[c]: the v bias rides through the attention (softmax rows sum to 1) into the projection bias.
__global__ void pad_proj_kernel(const __nv_bfloat16* __restrict__ w, const float* __restrict__ bp, const float* __restrict__ bv,
                                __nv_bfloat16* __restrict__ out, float* __restrict__ bp_eff, int C, int nH, int hd, int hdp) {
    const int n = blockIdx.x;                 // one CTA per output channel
    __shared__ float red[128];
    float acc = 0.f;
    for (int i = threadIdx.x; i < nH * hdp; i += blockDim.x) {
        const int d = i % hdp, h = i / hdp;
        const __nv_bfloat16 v = d < hd ? w[n * C + h * hd + d] : __float2bfloat16_rn(0.f);
        out[n * nH * hdp + i] = v;
        if (d < hd) acc = fmaf(__bfloat162float(v), bv[h * hd + d], acc);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) bp_eff[n] = bp[n] + red[0];
}
int attn_block_pad_proj(const __nv_bfloat16* w, const float* bp, const float* bv, __nv_bfloat16* out, float* bp_eff, int C, int nH, cudaStream_t s) {
    const int hd = C / nH;
    pad_proj_kernel<<<C, 128, 0, s>>>(w, bp, bv, out, bp_eff, C, nH, hd, 32);
    return check_cuda(cudaGetLastError(), "pad_proj launch");
}

// Host packing of the padded q/k/v weights of one block from the reference tensors (attn.qkv.weight [3C, C], attn.qkv.bias [3C],
// relative_position_bias_table [225, nH]); q rows carry head_dim^-0.5 (htsat.py:295,331) * log2(e) (the softmax runs in base 2).
int attn_block_pack(AttnBlockW& w, const std::vector<float>& qkv_w, const std::vector<float>& qkv_b, const std::vector<float>& rpb, int C, int nH) {
    const int hd = C / nH;
    if (C != AB_C || nH != AB_NH || hd != AB_HD) return set_error(ARD_ERR_SHAPE, "attn_block: built for C=96, 4 heads of 24");
    const float qs = AB_LOG2E / sqrtf((float)hd);
    std::vector<float> wqk((size_t)256 * C, 0.f), bq(128, 0.f), wv((size_t)128 * C, 0.f), bv(C, 0.f), tab((size_t)nH * AB_TAB_N, 0.f);
    for (int h = 0; h < nH; ++h)
        for (int d = 0; d < hd; ++d) {
            const int src = h * hd + d, dst = h * 32 + d;
            for (int k = 0; k < C; ++k) {
                wqk[(size_t)dst * C + k] = qkv_w[(size_t)src * C + k] * qs;
                wqk[(size_t)(128 + dst) * C + k] = qkv_w[(size_t)(C + src) * C + k];
                wv[(size_t)dst * C + k] = qkv_w[(size_t)(2 * C + src) * C + k];
            }
            bq[dst] = qkv_b[src] * qs;     // the k bias cancels in the softmax; the v bias (un-padded) is folded into the projection bias
            bv[src] = qkv_b[2 * C + src];
        }
    for (int h = 0; h < nH; ++h)
        for (int dy = 0; dy < 15; ++dy)
            for (int dx = 0; dx < 15; ++dx)
                tab[(size_t)h * AB_TAB_N + dy * AB_TAB_STRIDE + dx] = rpb[(size_t)(dy * 15 + dx) * nH + h] * AB_LOG2E;
    ARD_TRY(upload_bf16(w.wqk, wqk));
    ARD_TRY(upload_f32(w.bq, bq));
    ARD_TRY(upload_bf16(w.wv, wv));
    ARD_TRY(upload_f32(w.bv, bv));
    ARD_TRY(upload_f32(w.table, tab));
    w.ready = true;
    return 0;
}

// y = x + proj'(window_attention(LayerNorm(x))) for B clips of R x R tokens, C = 96. wp_pad [96, 128] bf16, bp [96] fp32 = the
// effective bias attn_block_pad_proj produced (device).
int attn_block_96(const float* x, float* out, const AttnBlockW& w, const __nv_bfloat16* wp_pad, const float* bp, const float* gamma,
                  const float* beta, int B, int R, int shift, int num_sms, cudaStream_t stream) {
    if (B <= 0) return 0;
    if (!w.ready) return set_error(ARD_ERR_STATE, "attn_block: weights not packed");
    if (R % 16 != 0) return set_error(ARD_ERR_SHAPE, "attn_block: needs an even number of windows per row (R %% 16 == 0), got R=%d", R);
    const long long windows = (long long)B * (R / 8) * (R / 8);
    if (windows / 2 > 0x3fffffffLL) return set_error(ARD_ERR_SHAPE, "attn_block: batch too large");
    CUtensorMap tqk, tv, tp;
    ARD_TRY(make_tmap_2d(&tqk, w.wqk.p, 2, AB_C, 256, (uint64_t)AB_C * 2, 32, 256, 64));
    ARD_TRY(make_tmap_2d(&tv, w.wv.p, 2, AB_C, 128, (uint64_t)AB_C * 2, 32, 128, 64));
    ARD_TRY(make_tmap_2d(&tp, wp_pad, 2, 128, AB_C, (uint64_t)128 * 2, 32, AB_C, 64));
    static bool attr_set = false;
    if (!attr_set) {
        ARD_CUDA(cudaFuncSetAttribute(attn_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES));
        attr_set = true;
    }
    AttnBlockParams p;
    p.x = x; p.out = out; p.bq = w.bq.as<float>(); p.bp = bp; p.gamma = gamma; p.beta = beta;
    p.table = w.table.as<float>(); p.R = R; p.shift = R > 8 ? shift : 0; p.n_tiles = (int)(windows / 2);
    const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
    const double M = (double)B * R * R;
    // executed tensor work: QK (N 256) + V^T (N 128) K 96, S (128x128x32) + PV (128x32x128) per head, proj K 128 N 96, per 128 tokens
    ProfScope ps(PROF_ATTN, stream, M * 2.0 * (384.0 * 96 + 4.0 * (128.0 * 32 + 32.0 * 128) + 128.0 * 96), M * AB_C * 8.0);
    ARD_CUDA(enqueue_pdl(attn_block_kernel, dim3(grid), dim3(AB_THREADS), AB_SMEM_BYTES, stream, tqk, tv, tp, p));
    return check_cuda(cudaGetLastError(), "attn_block launch");
}

}  // namespace ard

#ifdef ARD_AB_TRACE
extern "C" int ard_debug_ab_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, ard::g_ab_trace, sizeof(long long) * 4 * 32 * 24);
}
#endif
