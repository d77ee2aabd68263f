// PCA sufficient statistics on the tensor cores:  sum[D] += sum_r x[r,:],   sumsq[D,D] += x^T x   (float64 accumulators).
//
// Reference: IncrementalPCA.partial_fit called per batch by compute_pca_components (src/residual.py:103-159) and by run_PCA
// (src/analyze_attention.py:13-59); the exact covariance is what those approximate. X^T X is a dense
// contraction over the sample axis (16.1 GFLOP per clip for the 4096-d attention maps), so it runs on the tcgen05 GEMM:
//   1. split_transpose: x (fp32, row stride ldx) -> P = [hi | hi | 2 lo]  (bf16, [D, 3 Rp], sample axis contiguous), with
//      hi = bf16(x), lo = bf16(x - hi): x = hi + lo to ~2^-17 relative. Column sums (fp64) are taken in the same pass.
//   2. G = [hi | hi] [hi | 2 lo]^T = hi^T hi + 2 hi^T lo     (one gemm_tc launch, both operands are windows of P; fp32 accumulation
//      in TMEM over 2 Rp terms)
//   3. fold: sumsq += (G + G^T) / 2 = hi^T hi + hi^T lo + lo^T hi   (fp64; the dropped lo^T lo term is ~2^-18 relative)
// Samples are processed in chunks so the bf16 scratch stays bounded.
#include "ard_handle.h"

namespace ard {

// x [rows, ldx] fp32 -> P [D, ldk] bf16, three planes of this chunk's R rows at k-offsets 0 (hi), Rp (hi), 2 Rp (2 lo): the GEMM reads
// A = P[:, 0 : 2 Rp] = [hi | hi] and W = P[:, Rp : 3 Rp] = [hi | 2 lo] out of the same buffer (12 B written per element instead of 16).
// Each CTA walks `slabs` consecutive 64-row slabs of one 64-column strip and keeps the column sums in registers: one fp64
// atomic per column per CTA (the per-slab atomics of the first version were 8 M per step).
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ x, long long ldx, long long r0, int R, int Rp, int D,
                                                             __nv_bfloat16* __restrict__ P, long long ldk, double* __restrict__ sum, int slabs) {
    __shared__ float tile[64][65];
    __shared__ float csum[4][64];
    const int db = blockIdx.y * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 x 4
    float colacc = 0.f;                                        // column db + tx, rows ty, ty + 4, ... of every slab
    for (int sl = 0; sl < slabs; ++sl) {
        const int rb = (blockIdx.x * slabs + sl) * 64;
        if (rb >= Rp) break;
        for (int rr = ty; rr < 64; rr += 4) {
            const int r = rb + rr, d = db + tx;
            const float v = (r < R && d < D) ? x[(r0 + r) * ldx + d] : 0.f;
            tile[rr][tx] = v;
            colacc += v;
        }
        __syncthreads();
        for (int dd = ty; dd < 64; dd += 4) {
            const int d = db + dd, r = rb + tx;
            const float v = tile[tx][dd];
            if (d < D && r < Rp) {
                const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                const __nv_bfloat16 lo2 = __float2bfloat16_rn(2.0f * (v - __bfloat162float(hi)));
                P[(long long)d * ldk + r] = hi;
                P[(long long)d * ldk + Rp + r] = hi;
                P[(long long)d * ldk + 2 * Rp + r] = lo2;
            }
        }
        __syncthreads();
    }
    csum[ty][tx] = colacc;
    __syncthreads();
    if (ty == 0 && db + tx < D) atomicAdd(sum + db + tx, (double)csum[0][tx] + (double)csum[1][tx] + (double)csum[2][tx] + (double)csum[3][tx]);
}

// sumsq[i,j] += 0.5 * (G[i,j] + G[j,i])   (single partial: the 4096-d attention-map statistics, HBM-bound on the fp64 accumulator)
__global__ void __launch_bounds__(256) stats_fold1_kernel(const float* __restrict__ G, int D, double* __restrict__ sumsq) {
    __shared__ float tt[32][33];
    const int ib = blockIdx.y * 32, jb = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int j = jb + k, i = ib + tx;
        tt[k][tx] = (i < D && j < D) ? G[(long long)j * D + i] : 0.f;    // G[j, i]
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int i = ib + k, j = jb + tx;
        if (i < D && j < D) {
            const long long o = (long long)i * D + j;
            sumsq[o] += 0.5 * ((double)G[o] + (double)tt[tx][k]);
        }
    }
}

// sumsq[i,j] += 0.5 * (Gs[i,j] + Gs[j,i]),  Gs = sum over the S split-K partials (partial s = rows [s * srows, +D) of G), in fp64
__global__ void __launch_bounds__(256) stats_fold_kernel(const float* __restrict__ G, int D, int S, long long srows, double* __restrict__ sumsq) {
    __shared__ double tt[32][33];
    const int ib = blockIdx.y * 32, jb = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int j = jb + k, i = ib + tx;
        double a = 0.0;
        if (i < D && j < D)
            for (int sp = 0; sp < S; ++sp) a += (double)G[(sp * srows + j) * D + i];    // Gs[j, i]
        tt[k][tx] = a;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int i = ib + k, j = jb + tx;
        if (i < D && j < D) {
            double a = 0.0;
            for (int sp = 0; sp < S; ++sp) a += (double)G[(sp * srows + i) * D + j];
            sumsq[(long long)i * D + j] += 0.5 * (a + tt[tx][k]);
        }
    }
}

static DevBuf g_p, g_g;   // scratch shared by all calls on this process's device (one process per GPU)

int stats_accumulate(const float* x, long long rows, long long ldx, int D, double* sum, double* sumsq, cudaStream_t s) {
    if (rows <= 0) return 0;
    if (D <= 0 || ldx < D) return set_error(ARD_ERR_SHAPE, "stats_accumulate: bad D=%d ldx=%lld", D, ldx);
    if (D % 4) return set_error(ARD_ERR_SHAPE, "stats_accumulate: D=%d must be a multiple of 4", D);
    int dev = 0, sms = 148;
    ARD_CUDA(cudaGetDevice(&dev));
    ARD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // chunk of samples: bounded scratch (<= 2 x 256 MB of bf16) and a long enough K for the GEMM to amortise its [D,D] output
    long long chunk = (64LL << 20) / D;
    chunk = chunk < 1024 ? 1024 : chunk;
    chunk = (chunk / 64) * 64;
    if (chunk > rows) chunk = ((rows + 7) / 8) * 8;
    const long long ldk = 3 * chunk;
    ARD_TRY(g_p.ensure((size_t)D * ldk * 2));
    // Few output tiles and a long sample axis (the per-layer residual moments: D = 96..768, up to a million rows per call): split
    // the K range over enough CTAs to fill the GPU; the partial [D, D] products are summed in fp64 by the fold.
    const int mn_tiles = ((D + 127) / 128) * ((D + 127) / 128);
    const long long srows = ((D + 255) / 256) * 256LL;
    int splitk = 1;
    if (mn_tiles < sms) {
        splitk = (2 * sms + mn_tiles - 1) / mn_tiles;
        const long long max_by_k = (2 * (chunk < rows ? chunk : rows) / 64) / 8;    // keep >= 8 k-blocks per split
        if (splitk > max_by_k) splitk = (int)(max_by_k < 1 ? 1 : max_by_k);
    }
    ARD_TRY(g_g.ensure((size_t)(splitk > 1 ? splitk * srows : D) * D * 4));
    for (long long r0 = 0; r0 < rows; r0 += chunk) {
        const int R = (int)((rows - r0) < chunk ? (rows - r0) : chunk);
        const int Rp = (R + 7) & ~7;
        {
            ProfScope ps(PROF_OTHER, s, 0.0, (double)R * D * 12.0);
            const int nslab = (Rp + 63) / 64;
            const int slabs = nslab >= 64 ? 8 : (nslab >= 8 ? 4 : 1);     // rows per CTA: fewer atomics, still >= 148 CTAs for the big shapes
            dim3 grid((unsigned)((nslab + slabs - 1) / slabs), (unsigned)((D + 63) / 64));
            split_transpose_kernel<<<grid, 256, 0, s>>>(x, ldx, r0, R, Rp, D, g_p.as<__nv_bfloat16>(), ldk, sum, slabs);
            ARD_TRY(check_cuda(cudaGetLastError(), "stats split_transpose launch"));
        }
        GemmArgs g;
        g.A = g_p.as<__nv_bfloat16>(); g.lda = ldk; g.W = g_p.as<__nv_bfloat16>() + Rp; g.ldw = ldk; g.out = g_g.as<float>(); g.ldo = D;
        g.M = D; g.N = D; g.K = 2 * Rp;
        if (D >= 1024 && Rp >= 512) { g.force_pair = 1; g.force_bn = 192; }   // 256 x 192 CTA-pair tiles: the 128 x 128 fp32-output tile is
                                                                                 // L2-operand bound (32 KB of operands per 256 tensor cycles)
        int S = 1;
        if (splitk > 1) {
            g.splitk = splitk; g.split_rows = (int)srows; g.force_pair = -1; g.force_bn = 128;
            S = gemm_splitk_used(g.K, splitk);
        }
        ARD_TRY(gemm_bf16(g, sms, s));
        {
            ProfScope ps(PROF_OTHER, s, 0.0, (double)D * D * (20.0 + 4.0 * S));
            dim3 grid((unsigned)((D + 31) / 32), (unsigned)((D + 31) / 32));
            if (S == 1) stats_fold1_kernel<<<grid, 256, 0, s>>>(g_g.as<float>(), D, sumsq);
            else stats_fold_kernel<<<grid, 256, 0, s>>>(g_g.as<float>(), D, S, srows, sumsq);
            ARD_TRY(check_cuda(cudaGetLastError(), "stats fold launch"));
        }
    }
    return 0;
}

}  // namespace ard
