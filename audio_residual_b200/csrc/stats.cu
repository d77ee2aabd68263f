// PCA sufficient statistics on the tensor cores:  sum[D] += sum_r x[r,:],   sumsq[D,D] += x^T x   (float64 accumulators).
//
// Reference: IncrementalPCA.partial_fit called per batch by compute_pca_components (src/residual.py:103-159) and by run_PCA
// (src/analyze_attention.py:13-59); the exact covariance is what those approximate. X^T X is a dense
// contraction over the sample axis (16.1 GFLOP per clip for the 4096-d attention maps), so it runs on the tcgen05 GEMM:
//   1. split_transpose: x (fp32, row stride ldx) -> At = [hi | hi], Wt = [hi | 2 lo]  (bf16, [D, 2 Rp], sample axis contiguous),
//      with hi = bf16(x), lo = bf16(x - hi): x = hi + lo to ~2^-17 relative. Column sums (fp64) are taken in the same pass.
//   2. G = At Wt^T = hi^T hi + 2 hi^T lo                     (one gemm_tc launch, fp32 accumulation in TMEM over 2 Rp terms)
//   3. fold: sumsq += (G + G^T) / 2 = hi^T hi + hi^T lo + lo^T hi   (fp64; the dropped lo^T lo term is ~2^-18 relative)
// Samples are processed in chunks so the bf16 scratch stays bounded.
#include "ard_handle.h"

namespace ard {

// x [rows, ldx] fp32 (columns c0..c0+D) -> At/Wt [D, ldk] bf16 with this chunk's R rows at k-offsets [0,R) and [Rp, Rp+R)
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ x, long long ldx, long long r0, int R, int Rp, int D,
                                                             __nv_bfloat16* __restrict__ At, __nv_bfloat16* __restrict__ Wt, long long ldk,
                                                             double* __restrict__ sum) {
    __shared__ float tile[64][65];
    const int rb = blockIdx.x * 64, db = blockIdx.y * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 x 4
    for (int rr = ty; rr < 64; rr += 4) {
        const int r = rb + rr, d = db + tx;
        tile[rr][tx] = (r < R && d < D) ? x[(r0 + r) * ldx + d] : 0.f;
    }
    __syncthreads();
    for (int dd = ty; dd < 64; dd += 4) {
        const int d = db + dd, r = rb + tx;
        const float v = tile[tx][dd];
        // column sum of this 64-row slab: reduce over the 64 lanes holding the slab's rows (two warps), fp64 atomics
        float cs = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
        if ((tx & 31) == 0 && d < D) atomicAdd(sum + d, (double)cs);
        if (d < D && r < Rp) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            const __nv_bfloat16 lo2 = __float2bfloat16_rn(2.0f * (v - __bfloat162float(hi)));
            At[(long long)d * ldk + r] = hi;
            At[(long long)d * ldk + Rp + r] = hi;
            Wt[(long long)d * ldk + r] = hi;
            Wt[(long long)d * ldk + Rp + r] = lo2;
        }
    }
}

// sumsq[i,j] += 0.5 * (G[i,j] + G[j,i])
__global__ void __launch_bounds__(256) stats_fold_kernel(const float* __restrict__ G, int D, double* __restrict__ sumsq) {
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

static DevBuf g_at, g_wt, g_g;   // scratch shared by all calls on this process's device (one process per GPU)

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
    const long long ldk = 2 * chunk;
    ARD_TRY(g_at.ensure((size_t)D * ldk * 2));
    ARD_TRY(g_wt.ensure((size_t)D * ldk * 2));
    ARD_TRY(g_g.ensure((size_t)D * D * 4));
    for (long long r0 = 0; r0 < rows; r0 += chunk) {
        const int R = (int)((rows - r0) < chunk ? (rows - r0) : chunk);
        const int Rp = (R + 7) & ~7;
        {
            ProfScope ps(PROF_OTHER, s, 0.0, (double)R * D * 12.0);
            dim3 grid((unsigned)((Rp + 63) / 64), (unsigned)((D + 63) / 64));
            split_transpose_kernel<<<grid, 256, 0, s>>>(x, ldx, r0, R, Rp, D, g_at.as<__nv_bfloat16>(), g_wt.as<__nv_bfloat16>(), ldk, sum);
            ARD_TRY(check_cuda(cudaGetLastError(), "stats split_transpose launch"));
        }
        GemmArgs g;
        g.A = g_at.as<__nv_bfloat16>(); g.lda = ldk; g.W = g_wt.as<__nv_bfloat16>(); g.ldw = ldk; g.out = g_g.as<float>(); g.ldo = D;
        g.M = D; g.N = D; g.K = 2 * Rp;
        ARD_TRY(gemm_bf16(g, sms, s));
        {
            ProfScope ps(PROF_OTHER, s, 0.0, (double)D * D * 24.0);
            dim3 grid((unsigned)((D + 31) / 32), (unsigned)((D + 31) / 32));
            stats_fold_kernel<<<grid, 256, 0, s>>>(g_g.as<float>(), D, sumsq);
            ARD_TRY(check_cuda(cudaGetLastError(), "stats fold launch"));
        }
    }
    return 0;
}

}  // namespace ard
