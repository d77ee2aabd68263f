// Small fp32 kernels around the encoder: audio_projection + L2 normalise (clap_module/model.py:539-543, :739-741),
// the ResiDual fold (src/residual.py:29-42 composed with WindowAttention.proj htsat.py:355), the token-semantic head
// (htsat.py:797-821) and the PCA moment accumulation (replaces IncrementalPCA.partial_fit, src/residual.py:137-138).
#include "ard_common.cuh"
#include "ard_internal.h"

namespace ard {

// ------------------------------------------------------------------------------------------------ small-batch Linear
// y[b, n] = act(sum_k x[b,k] W[n,k] + bias[n]);  one CTA = CL rows of x (held in smem) x a slice of the N outputs;
// one warp per output n, lanes stride over k (coalesced reads of W rows), CL accumulators per lane.
constexpr int LS_CL = 8;
__global__ void __launch_bounds__(256) linear_small_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W,
                                                          const float* __restrict__ bias, float* __restrict__ y, int ldy, int B, int N,
                                                          int K, int act, int n_per_cta) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float xs[];   // [LS_CL][K]
    const int b0 = blockIdx.y * LS_CL;
    const int nb = min(LS_CL, B - b0);
    for (int i = threadIdx.x; i < LS_CL * K; i += blockDim.x) {
        const int r = i / K, k = i - r * K;
        xs[i] = r < nb ? x[(long long)(b0 + r) * ldx + k] : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_begin = blockIdx.x * n_per_cta;
    const int n_end = min(n_begin + n_per_cta, N);
    for (int n = n_begin + warp; n < n_end; n += 8) {
        float acc[LS_CL];
#pragma unroll
        for (int r = 0; r < LS_CL; ++r) acc[r] = 0.f;
        const float* wr = W + (long long)n * K;
        if ((K & 3) == 0) {                           // 16-byte loads of the weight row and of the staged activations
            for (int k = lane * 4; k < K; k += 128) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
                for (int r = 0; r < LS_CL; ++r) {
                    const float4 xv = *reinterpret_cast<const float4*>(xs + r * K + k);
                    acc[r] = fmaf(w.x, xv.x, fmaf(w.y, xv.y, fmaf(w.z, xv.z, fmaf(w.w, xv.w, acc[r]))));
                }
            }
        } else {
            for (int k = lane; k < K; k += 32) {
                const float w = __ldg(wr + k);
#pragma unroll
                for (int r = 0; r < LS_CL; ++r) acc[r] = fmaf(w, xs[r * K + k], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < LS_CL; ++r)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
        if (lane == 0) {
            const float bv = bias ? bias[n] : 0.f;
            for (int r = 0; r < nb; ++r) {
                float v = acc[r] + bv;
                if (act == ARD_ACT_RELU) v = fmaxf(v, 0.f);
                y[(long long)(b0 + r) * ldy + n] = v;
            }
        }
    }
}

int linear_small(const float* x, int ldx, const float* W, const float* bias, float* y, int ldy, int B, int N, int K, int act,
                 cudaStream_t s) {
    if (B <= 0) return 0;
    const int smem = LS_CL * K * 4;
    if (smem > 48 * 1024) return set_error(ARD_ERR_SHAPE, "linear_small: K=%d too large", K);
    const int n_per_cta = 64;
    dim3 grid((N + n_per_cta - 1) / n_per_cta, (B + LS_CL - 1) / LS_CL);
    ProfScope ps(PROF_HEAD, s, 2.0 * B * N * K, 4.0 * ((double)N * K + (double)B * K + (double)B * N));
    ARD_CUDA(enqueue_pdl(linear_small_kernel, grid, dim3(256), smem, s, x, ldx, W, bias, y, ldy, B, N, K, act, n_per_cta));
    return check_cuda(cudaGetLastError(), "linear_small launch");
}

// F.normalize(x, dim=-1): x / max(||x||_2, 1e-12); one warp per row
__global__ void l2_normalize_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int N) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    float s = 0.f;
    for (int k = lane; k < N; k += 32) {
        const float v = x[(long long)row * N + k];
        s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
    for (int k = lane; k < N; k += 32) y[(long long)row * N + k] = x[(long long)row * N + k] * inv;
}

int l2_normalize(const float* x, float* y, int B, int N, cudaStream_t s) {
    if (B <= 0) return 0;
    ARD_CUDA(enqueue_pdl(l2_normalize_kernel, dim3((B + 7) / 8), dim3(256), 0, s, x, y, B, N));
    return check_cuda(cudaGetLastError(), "l2_normalize launch");
}

// backward of F.normalize: y = p / n, n = max(||p||, 1e-12):  dp = (g - y (y . g)) / n
__global__ void l2_normalize_bwd_kernel(const float* __restrict__ p, const float* __restrict__ g, float* __restrict__ out, int B, int N) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    float s = 0.f, d = 0.f;
    for (int k = lane; k < N; k += 32) {
        const float v = p[(long long)row * N + k];
        s = fmaf(v, v, s);
        d = fmaf(v, g[(long long)row * N + k], d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
    const float yg = d * inv;   // y . g
    for (int k = lane; k < N; k += 32) {
        const long long i = (long long)row * N + k;
        out[i] = (g[i] - p[i] * inv * yg) * inv;
    }
}
int l2_normalize_bwd(const float* p, const float* g, float* out, int B, int N, cudaStream_t s) {
    if (B <= 0) return 0;
    l2_normalize_bwd_kernel<<<(B + 7) / 8, 256, 0, s>>>(p, g, out, B, N);
    return check_cuda(cudaGetLastError(), "l2_normalize_bwd launch");
}

// g *= (act > 0)   (ReLU backward from the saved post-activation)
__global__ void relu_bwd_mul_kernel(float* __restrict__ g, const float* __restrict__ act, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        if (!(act[i] > 0.f)) g[i] = 0.f;
}
int relu_bwd_mul(float* g, const float* act, long long n, cudaStream_t s) {
    if (n <= 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    relu_bwd_mul_kernel<<<(unsigned)blocks, 256, 0, s>>>(g, act, n);
    return check_cuda(cudaGetLastError(), "relu_bwd_mul launch");
}

// ------------------------------------------------------------------------------------------------ fp32 "TN" GEMM
// out[i][j] = sum_k a[k*lda + i] * (scale ? scale[k] : 1) * b[k*ldb + j];  i < Ma, j < Nb. 64x64 tile, 4x4 per thread.
// Only used to re-derive the folded ResiDual projection when lambda changes (C^3 flops, off the per-clip path).
template <bool OUT_BF16>
__global__ void __launch_bounds__(256) sgemm_tn_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                                      const float* __restrict__ scale, void* __restrict__ out, int ldo, int Ma, int Nb,
                                                      int K) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int t = threadIdx.x; t < 16 * 64; t += 256) {
            const int kk = t >> 6, c = t & 63;
            const int k = k0 + kk;
            float av = 0.f, bv = 0.f;
            if (k < K) {
                if (i0 + c < Ma) av = a[(long long)k * lda + i0 + c] * (scale ? scale[k] : 1.0f);
                if (j0 + c < Nb) bv = b[(long long)k * ldb + j0 + c];
            }
            As[kk][c] = av;
            Bs[kk][c] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float ar[4], br[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                ar[r] = As[kk][ty * 4 + r];
                br[r] = Bs[kk][tx * 4 + r];
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(ar[r], br[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = i0 + ty * 4 + r, j = j0 + tx * 4 + c;
            if (i < Ma && j < Nb) {
                if constexpr (OUT_BF16)
                    reinterpret_cast<__nv_bfloat16*>(out)[(long long)i * ldo + j] = __float2bfloat16_rn(acc[r][c]);
                else
                    reinterpret_cast<float*>(out)[(long long)i * ldo + j] = acc[r][c];
            }
        }
}

static int sgemm_tn(const float* a, int lda, const float* b, int ldb, const float* scale, void* out, int ldo, bool out_bf16, int Ma, int Nb,
                    int K, cudaStream_t s) {
    dim3 grid((Nb + 63) / 64, (Ma + 63) / 64);
    ProfScope ps(PROF_OTHER, s, 2.0 * Ma * Nb * K, 4.0 * ((double)Ma * K + (double)Nb * K + (double)Ma * Nb));
    if (out_bf16)
        sgemm_tn_kernel<true><<<grid, 256, 0, s>>>(a, lda, b, ldb, scale, out, ldo, Ma, Nb, K);
    else
        sgemm_tn_kernel<false><<<grid, 256, 0, s>>>(a, lda, b, ldb, scale, out, ldo, Ma, Nb, K);
    return check_cuda(cudaGetLastError(), "sgemm_tn launch");
}

// ResiDual (src/residual.py:37-42) applied to y = x Wp^T + bp:   r = ((y - mu) B^T * lambda) B = x (M Wp)^T + (bp - mu) M
// with M = B^T diag(lambda) B (symmetric).  Inputs: proj_w [C,C] fp32, dmean = bp - mu [C], basis [K,C], lam [K].
int residual_matrix(const float* basis, const float* lam, int C, int K, float* M, cudaStream_t s) {
    return sgemm_tn(basis, C, basis, C, lam, M, C, false, C, C, K, s);               // M[i][j] = sum_k B[k][i] lam[k] B[k][j]
}
int residual_fold(const float* proj_w, const float* dmean, const float* basis, const float* lam, int C, int K, float* Mtmp,
                  __nv_bfloat16* w_out, float* b_out, cudaStream_t s, float* w_out_f32) {
    ARD_TRY(residual_matrix(basis, lam, C, K, Mtmp, s));
    if (w_out_f32 != nullptr) {   // keep the fp32 fold too (the fp32-grade mode splits it into bf16 terms); bf16(acc) is unchanged
        ARD_TRY(sgemm_tn(Mtmp, C, proj_w, C, nullptr, w_out_f32, C, false, C, C, C, s));
        ARD_TRY(f32_to_bf16(w_out_f32, w_out, (long long)C * C, 1.0f, s));
    } else {
        ARD_TRY(sgemm_tn(Mtmp, C, proj_w, C, nullptr, w_out, C, true, C, C, C, s));  // W'[i][j] = sum_c M[c][i] Wp[c][j]
    }
    ARD_TRY(sgemm_tn(dmean, 1, Mtmp, C, nullptr, b_out, C, false, 1, C, C, s));      // b'[j] = sum_c (bp-mu)[c] M[c][j]
    return 0;
}

// ------------------------------------------------------------------------------------------------ token-semantic head
// forward_features tail, htsat.py:798-821. normed [B, 64, C] (token = f*8 + t) is regrouped to x'[c][fb][g*8+t] with
// freq row f = g*2 + fb (4 time-quarters g stacked along frequency, 2 frequency bins each), then
//   fine_grained_embedding[b, 32*T' + r, c] = mean_fb x'[c][fb][T']                 (interpolate x32, utils.py:209-224)
//   y = Conv2d(C -> 527, kernel (2,3), padding (0,1))(x') -> [B, 527, 32]
//   framewise_output[b, 32*T' + r, o] = sigmoid(y[o][T']);  clipwise_output[b, o] = sigmoid(mean_T' y[o][T'])
// The conv is run as a GEMM over an im2col matrix A[(b,T'), (c, fb, kw)] (bf16) against tscam_conv.weight.view(527, 6C).
__global__ void tscam_im2col_kernel(const float* __restrict__ normed, __nv_bfloat16* __restrict__ A, int B, int C) {
    const long long total = (long long)B * 32 * C * 6;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int kw = (int)(i % 3);
        const int fb = (int)((i / 3) % 2);
        const int c = (int)((i / 6) % C);
        const long long row = i / (6LL * C);
        const int Tp = (int)(row % 32);
        const long long b = row / 32;
        const int tt = Tp + kw - 1;
        float v = 0.f;
        if (tt >= 0 && tt < 32) {
            const int gq = tt >> 3, t = tt & 7;
            const int f = gq * 2 + fb;
            v = normed[(b * 64 + f * 8 + t) * C + c];
        }
        A[i] = __float2bfloat16_rn(v);
    }
}

__global__ void tscam_finish_kernel(const float* __restrict__ y /*[B*32, ldy]*/, int ldy, float* __restrict__ framewise,
                                    float* __restrict__ clipwise, int B, int NC) {
    // one CTA per (clip); threads over classes
    const long long b = blockIdx.x;
    for (int o = threadIdx.x; o < NC; o += blockDim.x) {
        float s = 0.f;
        for (int Tp = 0; Tp < 32; ++Tp) {
            const float v = y[(b * 32 + Tp) * ldy + o];
            s += v;
            if (framewise) {
                const float sg = 1.0f / (1.0f + expf(-v));
                for (int r = 0; r < 32; ++r) framewise[(b * 1024 + Tp * 32 + r) * NC + o] = sg;
            }
        }
        if (clipwise) clipwise[b * NC + o] = 1.0f / (1.0f + expf(-s * (1.0f / 32.0f)));
    }
}

__global__ void fine_grained_kernel(const float* __restrict__ normed, float* __restrict__ fine, int B, int C) {
    const long long total = (long long)B * 1024 * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int row = (int)((i / C) % 1024);
        const long long b = i / (1024LL * C);
        const int Tp = row >> 5;
        const int gq = Tp >> 3, t = Tp & 7;
        const float v0 = normed[(b * 64 + (gq * 2 + 0) * 8 + t) * C + c];
        const float v1 = normed[(b * 64 + (gq * 2 + 1) * 8 + t) * C + c];
        fine[i] = (v0 + v1) * 0.5f;
    }
}

int tscam_im2col(const float* normed, __nv_bfloat16* A, int B, int C, cudaStream_t s) {
    tscam_im2col_kernel<<<148 * 8, 256, 0, s>>>(normed, A, B, C);
    return check_cuda(cudaGetLastError(), "tscam_im2col launch");
}
int tscam_finish(const float* y, int ldy, float* framewise, float* clipwise, int B, int NC, cudaStream_t s) {
    tscam_finish_kernel<<<B, 256, 0, s>>>(y, ldy, framewise, clipwise, B, NC);
    return check_cuda(cudaGetLastError(), "tscam_finish launch");
}
int fine_grained(const float* normed, float* fine, int B, int C, cudaStream_t s) {
    fine_grained_kernel<<<148 * 8, 256, 0, s>>>(normed, fine, B, C);
    return check_cuda(cudaGetLastError(), "fine_grained launch");
}

}  // namespace ard
