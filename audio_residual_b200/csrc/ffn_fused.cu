// Fused Swin FFN for the 96-channel stage:  out = x + fc2(GELU(fc1(LayerNorm(x))))  (+ optional second residual)
//
// Reference: SwinTransformerBlock.forward htsat.py:479-480 (`x = x + drop_path(mlp(norm2(x)))`, Mlp htsat.py:158-164) and
// the doubled form of the ResiDual-patched block, src/residual.py:93-96.
//
// Unfused, one FFN at stage 0 moves 34 bytes per token-channel through HBM (LN out, 4C hidden written and re-read) and is
// ~6x over its HBM floor; here x is read once (+ once more for the residual add, from L2) and the result written once.
// Both weight matrices (2 x 72 KB bf16) stay resident in shared memory for the life of the persistent CTA, the 4C-wide
// hidden activation only ever exists as a 128x64 TMEM accumulator and a 16 KB bf16 shared-memory operand tile.
//
//   warps 0-3   : LayerNorm producers (coalesced fp32 loads, stats by warp shuffle, bf16 rows written straight into the
//                 SWIZZLE_64B K-major A-operand layout tcgen05 expects) and, one tile later, the output epilogue
//                 (TMEM -> + b2 + residual(s) -> swizzled staging -> TMA store)
//   warp 4      : TMEM allocator, one-time TMA load of W1/W2, single-thread tcgen05.mma issue. Per 64-wide hidden chunk j:
//                 H_j = A1 W1_j^T (M128 N64 K96), then Y += GELU(H_j) W2[:, j]^T (M128 N96 K64), software-pipelined so
//                 the tensor core computes H_{j+1} while the GELU warps work on H_j
//   warps 5-20  : 16 GELU warps (4 per TMEM lane quadrant, 16 hidden columns each): tcgen05.ld H_j, + b1, exact-erf GELU,
//                 bf16 pack, write the A2 operand tile (SWIZZLE_128B). Four warps per scheduler hide the latency of the
//                 dependent FMA/MUFU chains; with fewer (8) the kernel ran at 49% issue utilisation.
//
// The kernel is bound by the GELU warps' CUDA-core issue rate (~16 instructions per hidden element), not by the tensor
// core (K = 96) and no longer by HBM.
#include <stdlib.h>

#include "ard_common.cuh"
#include "ard_internal.h"

namespace ard {

constexpr int FF_C = 96;
constexpr int FF_HD = 4 * FF_C;         // 384
constexpr int FF_NCH = FF_HD / 64;      // 6 hidden chunks of 64
constexpr int FF_BM = 128;
constexpr int FF_GELU_WARPS = 16;
constexpr int FF_THREADS = (5 + FF_GELU_WARPS) * 32;   // 4 LN/out + 1 MMA + 16 GELU warps = 672

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
constexpr int FF_A2_BYTES = FF_BM * 128;        // 16384 per buffer
constexpr int FF_CST_OFF = FF_A2_OFF + 2 * FF_A2_BYTES;    // 4 warps x 2 buffers x (32 rows x 64 B)
constexpr int FF_VEC_OFF = FF_CST_OFF + 4 * 4096;          // b1[384] b2[96] gamma[96] beta[96]
constexpr int FF_BAR_OFF = FF_VEC_OFF + (FF_HD + 3 * FF_C) * 4;
constexpr int FF_SMEM_BYTES = FF_BAR_OFF + 256 + 1024;

constexpr int FF_NHB = 3;       // H accumulators in flight (the MMA thread runs up to 3 hidden chunks ahead of the GELU warps)
constexpr int FF_TM_H = 0;      // TMEM columns: H0 @0, H1 @64, H2 @128, Y0 @256, Y1 @384
constexpr int FF_TM_Y = 256;

struct FfnParams {
    const float* x;        // [M, 96] fp32  (LayerNorm input and first residual)
    const float* resid2;   // [M, 96] fp32 or null
    float* out;            // [M, 96] fp32
    const float* gamma;    // norm2
    const float* beta;
    const float* b1;       // [384]
    const float* b2;       // [96]
    int M;
    int debug;   // bit0: skip residual loads, bit1: skip GELU math, bit2: skip LN loads (timing experiments only)
};

__global__ void __launch_bounds__(FF_THREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const FfnParams p) {
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
    uint64_t* h_full = bars + 3;    // [3]
    uint64_t* h_free = bars + 6;    // [3]
    uint64_t* a2_full = bars + 9;   // [2]
    uint64_t* a2_free = bars + 11;  // [2]
    uint64_t* y_full = bars + 13;   // [2]
    uint64_t* y_free = bars + 15;   // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = (p.M + FF_BM - 1) / FF_BM;

    for (int i = threadIdx.x; i < FF_HD; i += FF_THREADS) b1s[i] = p.b1[i];
    for (int i = threadIdx.x; i < FF_C; i += FF_THREADS) {
        b2s[i] = p.b2[i];
        gs[i] = p.gamma[i];
        bs[i] = p.beta[i];
    }
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmW1);
        tma_prefetch_desc(&tmW2);
        mbar_init(w_full, 1);
        mbar_init(a1_full, 4);
        mbar_init(a1_free, 1);
        for (int i = 0; i < FF_NHB; ++i) {
            mbar_init(&h_full[i], 1);
            mbar_init(&h_free[i], FF_GELU_WARPS);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a2_full[i], FF_GELU_WARPS);
            mbar_init(&a2_free[i], 1);
            mbar_init(&y_full[i], 1);
            mbar_init(&y_free[i], FF_GELU_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 4) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp < 4) {
        // ============================================================ LayerNorm producers + output epilogue
        uint8_t* a1 = smem + FF_A1_OFF;
        // LayerNorm thread mapping: 8 lanes per row (12 contiguous channels each), 4 rows per warp instruction. The row
        // reductions are 3-step butterflies inside the 8-lane group: 6 shuffles per FOUR rows instead of 10 per row.
        const int l8 = lane & 7, rsub = lane >> 3;
        int it = 0;
        auto prefetch_tile = [&](int t) {   // pull the tile's x (and second-residual) rows into L2 one tile ahead
            if (t >= num_tiles || (p.debug & 4)) return;
            const long long r = (long long)t * FF_BM + warp * 32 + lane;
            if (r >= p.M) return;
            const char* a = reinterpret_cast<const char*>(p.x + r * FF_C);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 128));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 256));
            if (p.resid2 != nullptr) {
                const char* c = reinterpret_cast<const char*>(p.resid2 + r * FF_C);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(c));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(c + 128));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(c + 256));
            }
        };
        prefetch_tile(blockIdx.x);
        for (int tile = blockIdx.x;; tile += gridDim.x, ++it) {
            const bool have = tile < num_tiles;
            if (have) {
                prefetch_tile(tile + gridDim.x);
                // ---- LayerNorm of rows [tile*128 + warp*32, +32) -> A1 (bf16, SWIZZLE_64B K-major, 3 k-blocks of 32 columns).
                // 4 batches of 2 row-groups (8 rows), software-pipelined: batch k+1's loads fly while batch k is normalised;
                // batch 0's latency overlaps the wait for the previous tile's fc1 MMAs (A1 is single-buffered).
                const long long row0 = (long long)tile * FF_BM + warp * 32;
                float4 v[2][2][3];
                auto load_batch = [&](int bi, float4 (&dst)[2][3]) {
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const long long row = row0 + (bi * 2 + g) * 4 + rsub;
                        if (row < p.M && !(p.debug & 4)) {
                            const float4* xr = reinterpret_cast<const float4*>(p.x + row * FF_C + l8 * 12);
                            dst[g][0] = __ldg(xr); dst[g][1] = __ldg(xr + 1); dst[g][2] = __ldg(xr + 2);
                        } else {
                            dst[g][0] = dst[g][1] = dst[g][2] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                };
                load_batch(0, v[0]);
                mbar_wait(a1_free, (it & 1) ^ 1);
#pragma unroll
                for (int bi = 0; bi < ((p.debug & 16) ? 0 : 4); ++bi) {
                    if (bi + 1 < 4) load_batch(bi + 1, v[(bi + 1) & 1]);
                    float4 (&cur)[2][3] = v[bi & 1];
                    float sm[2], qv[2];
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        sm[g] = 0.f;
#pragma unroll
                        for (int q = 0; q < 3; ++q) sm[g] += (cur[g][q].x + cur[g][q].y) + (cur[g][q].z + cur[g][q].w);
                    }
#pragma unroll
                    for (int sh = 4; sh > 0; sh >>= 1)
#pragma unroll
                        for (int g = 0; g < 2; ++g) sm[g] += __shfl_xor_sync(0xffffffffu, sm[g], sh);
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const float mean = sm[g] * (1.0f / FF_C);
                        qv[g] = 0.f;
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            cur[g][q].x -= mean; cur[g][q].y -= mean; cur[g][q].z -= mean; cur[g][q].w -= mean;
                            qv[g] += (cur[g][q].x * cur[g][q].x + cur[g][q].y * cur[g][q].y) + (cur[g][q].z * cur[g][q].z + cur[g][q].w * cur[g][q].w);
                        }
                    }
#pragma unroll
                    for (int sh = 4; sh > 0; sh >>= 1)
#pragma unroll
                        for (int g = 0; g < 2; ++g) qv[g] += __shfl_xor_sync(0xffffffffu, qv[g], sh);
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const float rstd = rsqrtf(qv[g] * (1.0f / FF_C) + 1e-5f);
                        const int rr = warp * 32 + (bi * 2 + g) * 4 + rsub;
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const int c0 = l8 * 12 + q * 4;                       // 4 channels = 8 bytes of bf16, inside one 16-byte unit
                            const float4 gm = *reinterpret_cast<const float4*>(gs + c0);
                            const float4 bt = *reinterpret_cast<const float4*>(bs + c0);
                            uint2 pk;
                            pk.x = pack_bf16x2(fmaf(cur[g][q].x * rstd, gm.x, bt.x), fmaf(cur[g][q].y * rstd, gm.y, bt.y));
                            pk.y = pack_bf16x2(fmaf(cur[g][q].z * rstd, gm.z, bt.z), fmaf(cur[g][q].w * rstd, gm.w, bt.w));
                            const int kb = c0 >> 5, cc = c0 & 31;
                            *reinterpret_cast<uint2*>(a1 + kb * FF_A1_KB + rr * 64 + ((((cc >> 3) ^ ((rr >> 1) & 3))) << 4) + (cc & 7) * 2) = pk;
                        }
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(a1_full);
            }
            if (!have) break;
        }
    } else if (warp == 4) {
        // ============================================================ weight load + MMA issue (one thread)
        if (lane == 0) {
            mbar_expect_tx(w_full, FF_W1_BYTES + FF_W2_BYTES);
            for (int kb = 0; kb < 3; ++kb)
                for (int hf = 0; hf < 2; ++hf)
                    tma_load_2d(smem + FF_W1_OFF + kb * FF_W1_KB + hf * 192 * 64, &tmW1, w_full, kb * 32, hf * 192);
            for (int j = 0; j < FF_NCH; ++j) tma_load_2d(smem + FF_W2_OFF + j * FF_W2_KB, &tmW2, w_full, j * 64, 0);
            mbar_wait(w_full, 0);
            constexpr uint32_t idesc1 = umma_idesc_bf16(FF_BM, 64);
            constexpr uint32_t idesc2 = umma_idesc_bf16(FF_BM, FF_C);
            const uint32_t sW1 = smem_u32(smem + FF_W1_OFF), sW2 = smem_u32(smem + FF_W2_OFF);
            const uint32_t sA1 = smem_u32(smem + FF_A1_OFF), sA2 = smem_u32(smem + FF_A2_OFF);
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int yb = it & 1;
                mbar_wait(a1_full, it & 1);
                tc_fence_after();
                // Each H buffer is used twice per tile (chunks k and k+3) -> n-th use has parity (j/3)&1; each A2 buffer three times.
                auto issue_g1 = [&](int j) {
                    const int hb = j % FF_NHB;
                    mbar_wait(&h_free[hb], ((j / FF_NHB) & 1) ^ 1);
                    tc_fence_after();
#pragma unroll
                    for (int kb = 0; kb < 3; ++kb) {
                        const uint64_t da = umma_desc_sw64(sA1 + kb * FF_A1_KB);
                        const uint64_t db = umma_desc_sw64(sW1 + kb * FF_W1_KB + j * 64 * 64);
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            umma_bf16_ss(tmem_base + FF_TM_H + hb * 64, da + 2 * ks, db + 2 * ks, idesc1, (kb | ks) != 0);
                    }
                    umma_commit(&h_full[hb]);
                    if (j == FF_NCH - 1) umma_commit(a1_free);
                };
                auto issue_g2 = [&](int j) {
                    const int b = j & 1;
                    mbar_wait(&a2_full[b], (it + (j >> 1)) & 1);
                    if (j == 0) mbar_wait(&y_free[yb], ((it >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(sA2 + b * FF_A2_BYTES);
                    const uint64_t db = umma_desc_sw128(sW2 + j * FF_W2_KB);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_bf16_ss(tmem_base + FF_TM_Y + yb * 128, da + 2 * ks, db + 2 * ks, idesc2, (j | ks) != 0);
                    umma_commit(&a2_free[b]);
                };
#pragma unroll
                for (int j = 0; j < FF_NHB; ++j) issue_g1(j);
#pragma unroll
                for (int j = 0; j < FF_NCH; ++j) {
                    issue_g2(j);
                    if (j + FF_NHB < FF_NCH) issue_g1(j + FF_NHB);
                }
                umma_commit(&y_full[yb]);
            }
        }
    } else {
        // ============================================================ GELU warps: H_j (TMEM) -> bf16 A2 operand tile
        const int ew = warp - 5;
        const int quad = warp & 3;
        const int cg = ew >> 2;           // 16-column group inside the 64-wide hidden chunk (0..3)
        const int row = quad * 32 + lane; // row inside the tile == TMEM lane
        // Output epilogue of tile `pt` (iteration `pit`): this warp owns 24 output columns (cg*24 ..) of its 32 rows.
        // y = acc + b2 + x (+ resid2), written straight from registers (96 contiguous bytes per thread).
        auto y_epilogue = [&](int pt, int pit) {
            const int yb = pit & 1;
            const long long grow = (long long)pt * FF_BM + row;
            const bool ok = grow < p.M;
            mbar_wait(&y_full[yb], (pit >> 1) & 1);
            tc_fence_after();
            uint32_t a[3][8];
#pragma unroll
            for (int q = 0; q < 3; ++q)
                tmem_ld_32x32b_x8(tmem_base + FF_TM_Y + yb * 128 + cg * 24 + q * 8 + ((uint32_t)(quad * 32) << 16), a[q]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&y_free[yb]);
            if (!ok) return;
            const float4* rp = reinterpret_cast<const float4*>(p.x + grow * FF_C + cg * 24);
            const float4* rq = reinterpret_cast<const float4*>(p.resid2 + grow * FF_C + cg * 24);
            float4* op = reinterpret_cast<float4*>(p.out + grow * FF_C + cg * 24);
            const bool use_r = !(p.debug & 1);
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const float4 b4 = *reinterpret_cast<const float4*>(b2s + cg * 24 + q * 4);
                float4 o;
                o.x = __uint_as_float(a[q >> 1][(q & 1) * 4 + 0]) + b4.x;
                o.y = __uint_as_float(a[q >> 1][(q & 1) * 4 + 1]) + b4.y;
                o.z = __uint_as_float(a[q >> 1][(q & 1) * 4 + 2]) + b4.z;
                o.w = __uint_as_float(a[q >> 1][(q & 1) * 4 + 3]) + b4.w;
                if (use_r) {
                    const float4 r4 = rp[q];
                    o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
                    if (p.resid2 != nullptr) {
                        const float4 s4 = rq[q];
                        o.x += s4.x; o.y += s4.y; o.z += s4.z; o.w += s4.w;
                    }
                }
                if (!(p.debug & 8)) op[q] = o;
            }
        };
        int it = 0;
        int prev_tile = -1;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
#pragma unroll
            for (int j = 0; j < FF_NCH; ++j) {
                if (j == 1 && prev_tile >= 0) y_epilogue(prev_tile, it - 1);   // previous tile's fc2 has retired by now
                const int b = j & 1;
                const int hb = j % FF_NHB;
                mbar_wait(&h_full[hb], (j / FF_NHB) & 1);
                tc_fence_after();
                uint32_t v[16];
                tmem_ld_32x32b_x16(tmem_base + FF_TM_H + hb * 64 + cg * 16 + ((uint32_t)(quad * 32) << 16), v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&h_free[hb]);
                const float* bb = b1s + j * 64 + cg * 16;
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bb + i);
                    float x0 = __uint_as_float(v[i]) + b4.x, x1 = __uint_as_float(v[i + 1]) + b4.y;
                    float x2 = __uint_as_float(v[i + 2]) + b4.z, x3 = __uint_as_float(v[i + 3]) + b4.w;
                    if (!(p.debug & 2)) { x0 = gelu_erf(x0); x1 = gelu_erf(x1); x2 = gelu_erf(x2); x3 = gelu_erf(x3); }
                    pk[i / 2] = pack_bf16x2(x0, x1);
                    pk[i / 2 + 1] = pack_bf16x2(x2, x3);
                }
                mbar_wait(&a2_free[b], ((it + (j >> 1)) & 1) ^ 1);   // fc2 MMAs that read the previous contents have retired
                uint8_t* rowp = smem + FF_A2_OFF + b * FF_A2_BYTES + row * 128;
                const int sw = row & 7;
                *reinterpret_cast<uint4*>(rowp + (((cg * 2 + 0) ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(rowp + (((cg * 2 + 1) ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a2_full[b]);
            }
            prev_tile = tile;
        }
        if (prev_tile >= 0) y_epilogue(prev_tile, it - 1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// x_out = x + fc2(gelu(fc1(LN(x)))) (+ resid2). x_out may alias x (each element is read by the warp that later writes it).
int ffn_fused_96(const float* x, const float* resid2, float* out, long long M, const float* gamma, const float* beta,
                 const __nv_bfloat16* w1, const float* b1, const __nv_bfloat16* w2, const float* b2, int num_sms, cudaStream_t stream) {
    if (M <= 0) return 0;
    if (M > 0x7fffffffLL) return set_error(ARD_ERR_SHAPE, "ffn_fused: too many rows");
    CUtensorMap t1, t2;
    ARD_TRY(make_tmap_2d(&t1, w1, 2, FF_C, FF_HD, (uint64_t)FF_C * 2, 32, 192, 64));
    ARD_TRY(make_tmap_2d(&t2, w2, 2, FF_HD, FF_C, (uint64_t)FF_HD * 2, 64, FF_C, 128));
    static bool attr_set = false;
    if (!attr_set) {
        ARD_CUDA(cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM_BYTES));
        attr_set = true;
    }
    FfnParams p;
    p.x = x; p.resid2 = resid2; p.out = out; p.gamma = gamma; p.beta = beta; p.b1 = b1; p.b2 = b2; p.M = (int)M;
    p.debug = getenv("ARD_FFN_DEBUG") ? atoi(getenv("ARD_FFN_DEBUG")) : 0;
    const int tiles = (int)((M + FF_BM - 1) / FF_BM);
    const int grid = tiles < num_sms ? tiles : num_sms;
    const double MC = (double)M * FF_C;
    ProfScope ps(PROF_FFN, stream, 2.0 * M * FF_C * FF_HD * 2.0, MC * 4.0 * (2.0 + (resid2 ? 1.0 : 0.0)) + 2.0 * 2.0 * FF_C * FF_HD);
    ffn_fused_kernel<<<grid, FF_THREADS, FF_SMEM_BYTES, stream>>>(t1, t2, p);
    return check_cuda(cudaGetLastError(), "ffn_fused launch");
}

}  // namespace ard
