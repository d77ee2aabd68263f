// Small kernels either side of the encoder, and their C-ABI entry points:
//   * standalone ResiDual module forward / backward                      (src/residual.py:29-42)
//   * classification head: logits, CrossEntropyLoss, their backward       (src/training.py:28-31, src/linear.py:23-45)
//   * evaluation reductions: argmax, top-k hits, confusion matrix         (src/evaluation.py:159-177)
//   * batched ragged featuriser: repeatpad / pad / repeat, int16 PCM      (hook.py:175-188, training/data.py:93-99, :466-496)
//   * per-head attention-output tap (the `attn @ v` temporary)            (htsat.py:354)
#include "ard_common.cuh"
#include "ard_handle.h"

namespace ard {

// ------------------------------------------------------------------------------------------------ strided fp32 GEMM (small)
// out[i*ldo + j] = alpha * sum_k a[k*sak + i*sai] * b[k*sbk + j*sbj] (+ out if accumulate). 64x64 tile, 4x4 per thread.
// Only for head-sized problems (B x 512 x 50) and per-call constants; the token-sized contractions run on gemm_tc.
__global__ void __launch_bounds__(256) sgemm_strided_kernel(const float* __restrict__ a, long long sak, long long sai,
                                                           const float* __restrict__ b, long long sbk, long long sbj, float* __restrict__ out,
                                                           int ldo, int Ma, int Nb, int K, float alpha) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int t = threadIdx.x; t < 16 * 64; t += 256) {
            // k fastest when the k stride is 1 (row-major A read along its rows), otherwise the i/j index is fastest
            int kk, c;
            if (sak == 1) { kk = t & 15; c = t >> 4; } else { kk = t >> 6; c = t & 63; }
            const int k = k0 + kk;
            As[kk][c] = (k < K && i0 + c < Ma) ? a[(long long)k * sak + (long long)(i0 + c) * sai] : 0.f;
            if (sbk == 1) { kk = t & 15; c = t >> 4; } else { kk = t >> 6; c = t & 63; }
            const int k2 = k0 + kk;
            Bs[kk][c] = (k2 < K && j0 + c < Nb) ? b[(long long)k2 * sbk + (long long)(j0 + c) * sbj] : 0.f;
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
            if (i < Ma && j < Nb) out[(long long)i * ldo + j] = alpha * acc[r][c];
        }
}

static int sgemm_strided(const float* a, long long sak, long long sai, const float* b, long long sbk, long long sbj, float* out, int ldo,
                         int Ma, int Nb, int K, float alpha, cudaStream_t s) {
    if (Ma <= 0 || Nb <= 0) return 0;
    dim3 grid((Nb + 63) / 64, (Ma + 63) / 64);
    ProfScope ps(PROF_HEAD, s, 2.0 * Ma * Nb * K, 4.0 * ((double)Ma * K + (double)Nb * K + (double)Ma * Nb));
    sgemm_strided_kernel<<<grid, 256, 0, s>>>(a, sak, sai, b, sbk, sbj, out, ldo, Ma, Nb, K, alpha);
    return check_cuda(cudaGetLastError(), "sgemm_strided launch");
}

// ------------------------------------------------------------------------------------------------ cross entropy
// One warp per row: loss_b = logsumexp(z_b) - z_b[label_b]; dlogits = (softmax(z_b) - onehot(label_b)) / B.
// The mean over rows is a single-CTA second pass (deterministic summation order: no float atomics).
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ z, const long long* __restrict__ labels, int B, int N,
                                                     float* __restrict__ row_loss, float* __restrict__ dz) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* zr = z + (long long)row * N;
    float m = -INFINITY;
    for (int j = lane; j < N; j += 32) m = fmaxf(m, zr[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) sum += expf(zr[j] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const long long lab = labels[row];
    const float lse = m + logf(sum);
    if (lane == 0) row_loss[row] = (lab >= 0 && lab < N) ? lse - zr[lab] : 0.f;
    if (dz != nullptr) {
        const float invB = 1.0f / (float)B, inv = 1.0f / sum;
        for (int j = lane; j < N; j += 32) {
            const float p = expf(zr[j] - m) * inv;
            dz[(long long)row * N + j] = (p - (j == lab ? 1.0f : 0.0f)) * invB;
        }
    }
}
__global__ void __launch_bounds__(256) mean_rows_kernel(const float* __restrict__ v, int B, float* __restrict__ out) {
    __shared__ double part[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < B; i += 256) acc += (double)v[i];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)(part[0] / (double)B);
}
// db[n] = sum_b dz[b, n]   (one thread per class, rows in order: deterministic)
__global__ void colsum_kernel(const float* __restrict__ dz, int B, int N, float* __restrict__ db) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += dz[(long long)b * N + n];
    db[n] = acc;
}

// ------------------------------------------------------------------------------------------------ evaluation reductions
// One warp per sample: argmax (first maximal index, torch.argmax), rank of the target under sklearn's top_k_accuracy_score
// ordering (stable ascending argsort reversed: among equal scores the HIGHER index ranks first), confusion-matrix count.
__global__ void __launch_bounds__(256) eval_metrics_kernel(const float* __restrict__ sc, const long long* __restrict__ targets, long long n, int C,
                                                          int k, unsigned long long* __restrict__ counts, unsigned long long* __restrict__ cm,
                                                          long long* __restrict__ preds) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float* r = sc + row * C;
    const long long t = targets[row];
    const bool tv = t >= 0 && t < C;
    const float st = tv ? r[t] : 0.f;
    float best = -INFINITY;
    int bi = C;
    int ahead = 0;
    for (int j = lane; j < C; j += 32) {
        const float v = r[j];
        if (v > best) { best = v; bi = j; }          // within a lane j ascends: keeps the first maximal index
        if (tv && (v > st || (v == st && j > t))) ++ahead;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        ahead += __shfl_xor_sync(0xffffffffu, ahead, o);
    }
    if (lane == 0) {
        if (preds) preds[row] = bi;
        if (tv) {
            if (bi == t) atomicAdd(counts + 0, 1ULL);
            if (ahead < k) atomicAdd(counts + 1, 1ULL);
            if (cm && bi < C) atomicAdd(cm + t * C + bi, 1ULL);
        }
    }
}

// ------------------------------------------------------------------------------------------------ ragged featuriser
// out[b, i] for i < max_len from clip b of length n (<= max_len):
//   repeatpad: i < n * floor(max_len / n) ? clip[i % n] : 0      (data.py:471-480: repeat, then F.pad with zeros)
//   pad      : i < n ? clip[i] : 0                               (data.py:481-487)
//   repeat   : clip[i % n]                                       (data.py:488-491: repeat int(max_len/n)+1 times, cut)
// Samples: fp32, or int16 PCM mapped through int16_to_float32 = (x / 32767.0).astype(float32) (data.py:93-94; the fp32
// division equals numpy's float64 division rounded once for all 65536 inputs, checked exhaustively in tests/test_cpu_host.py). `quantize`: float32_to_int16 then int16_to_float32 (hook.py:177-179) applied on the way.
template <bool PCM16>
__global__ void __launch_bounds__(256) fill_clips_kernel(const void* __restrict__ flat, const long long* __restrict__ offsets,
                                                        const int* __restrict__ lengths, int max_len, int mode, int quantize,
                                                        float* __restrict__ out) {
    const int b = blockIdx.y;
    const int n = lengths ? lengths[b] : max_len;                       // NULL offsets / lengths: dense [B, max_len] input
    const long long off = offsets ? offsets[b] : (long long)b * max_len;
    const int full = n > 0 ? (max_len / n) * n : 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max_len; i += gridDim.x * blockDim.x) {
        float v = 0.f;
        const bool live = n > 0 && (mode == 2 || (mode == 0 ? i < full : i < n));
        if (live) {
            const int j = i < n ? i : i % n;
            if constexpr (PCM16) v = (float)reinterpret_cast<const short*>(flat)[off + j] / 32767.0f;   // == the float64 division rounded once, for every int16
            else v = reinterpret_cast<const float*>(flat)[off + j];
            if (quantize) {
                v = fminf(fmaxf(v, -1.0f), 1.0f);
                v = truncf(v * 32767.0f) / 32767.0f;        // astype(int16): truncation toward zero
            }
        }
        out[(long long)b * max_len + i] = v;
    }
}

// ------------------------------------------------------------------------------------------------ per-head output tap
// ao [B*T, C] bf16 in token order (attention output before proj) -> tap [B*nW, nH, 64, hd] fp32 in the window order of a block
// with cyclic shift `shift` (htsat.py:452-460: roll by -shift then window_partition), i.e. the `attn @ v` temporary of
// WindowAttention.forward (htsat.py:354) before its transpose(1, 2).
__global__ void __launch_bounds__(256) head_tap_kernel(const __nv_bfloat16* __restrict__ ao, float* __restrict__ tap, int B, int R, int C,
                                                      int nH, int shift) {
    const int hd = C / nH, nWr = R / 8;
    const long long total = (long long)B * R * R * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % hd);
        long long r = i / hd;
        const int tok = (int)(r % 64); r /= 64;
        const int h = (int)(r % nH); r /= nH;
        const int win = (int)(r % (nWr * nWr));
        const long long b = r / (nWr * nWr);
        const int wy = win / nWr, wx = win % nWr;
        int y = wy * 8 + tok / 8 + shift, x = wx * 8 + tok % 8 + shift;   // rolled image (y', x') <- original (y' + shift, x' + shift)
        if (y >= R) y -= R;
        if (x >= R) x -= R;
        tap[i] = __bfloat162float(ao[((b * R + y) * R + x) * C + h * hd + d]);
    }
}

int head_output_tap(const __nv_bfloat16* ao, float* tap, int B, int R, int C, int nH, int shift, cudaStream_t s) {
    ProfScope ps(PROF_OTHER, s, 0.0, 6.0 * B * R * R * C);
    head_tap_kernel<<<148 * 8, 256, 0, s>>>(ao, tap, B, R, C, nH, R > 8 ? shift : 0);
    return check_cuda(cudaGetLastError(), "head_tap launch");
}

// ------------------------------------------------------------------------------------------------ standalone ResiDual
// Packs the per-call constants of the module from device fp32 tensors: basis [K,D] -> bf16 [Kp,D] (zero rows K..Kp) and its
// transpose bf16 [D,Kp]; c0[k] = -sum_c basis[k][c] mean[c] (zero for k >= K); lam padded to Kp.
__global__ void residual_pack_kernel(const float* __restrict__ basis, const float* __restrict__ mean, const float* __restrict__ lam, int K,
                                     int Kp, int D, __nv_bfloat16* __restrict__ bb, __nv_bfloat16* __restrict__ bbT, float* __restrict__ c0,
                                     float* __restrict__ lamp) {
    const int k = blockIdx.x;     // one CTA per (padded) component row
    __shared__ float red[256];
    float acc = 0.f;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        const float v = k < K ? basis[(long long)k * D + c] : 0.f;
        bb[(long long)k * D + c] = __float2bfloat16_rn(v);
        bbT[(long long)c * Kp + k] = __float2bfloat16_rn(v);
        acc = fmaf(v, mean[c], acc);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        c0[k] = -red[0];
        lamp[k] = k < K ? lam[k] : 0.f;
    }
}

struct ResidualScratch {
    DevBuf xb, gb, M, Mb, bias, bb, bbT, c0, lamp, coef, gcoef, gsc;
};
static ResidualScratch& rscratch() {
    static ResidualScratch r;
    return r;
}

static int device_sms() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace ard

using namespace ard;

extern "C" {

int ard_residual_forward(const float* x, const float* mean, const float* basis, const float* lam, float* out, long long rows, int D, int K,
                         void* stream) {
    if (!x || !mean || !basis || !lam || !out) return set_error(ARD_ERR_SHAPE, "ard_residual_forward: null argument");
    if (rows <= 0 || D <= 0 || K <= 0 || K > D || (D % 16) != 0) return set_error(ARD_ERR_SHAPE, "ard_residual_forward: bad shape rows=%lld D=%d K=%d", rows, D, K);
    cudaStream_t s = (cudaStream_t)stream;
    ResidualScratch& w = rscratch();
    ARD_TRY(w.xb.ensure((size_t)rows * D * 2));
    ARD_TRY(w.M.ensure((size_t)D * D * 4));
    ARD_TRY(w.Mb.ensure((size_t)D * D * 2));
    ARD_TRY(w.bias.ensure((size_t)D * 4));
    // M = B^T diag(lam) B  (symmetric), then  out = x M - mean M
    ARD_TRY(residual_matrix(basis, lam, D, K, w.M.as<float>(), s));
    ARD_TRY(f32_to_bf16(w.M.as<float>(), w.Mb.as<__nv_bfloat16>(), (long long)D * D, 1.0f, s));
    ARD_TRY(sgemm_strided(mean, 1, 0, w.M.as<float>(), D, 1, w.bias.as<float>(), D, 1, D, D, -1.0f, s));
    ARD_TRY(f32_to_bf16(x, w.xb.as<__nv_bfloat16>(), rows * D, 1.0f, s));
    GemmArgs g;
    g.A = w.xb.as<__nv_bfloat16>(); g.lda = D; g.W = w.Mb.as<__nv_bfloat16>(); g.ldw = D; g.out = out; g.ldo = D;
    g.M = (int)rows; g.N = D; g.K = D; g.bias = w.bias.as<float>();
    return gemm_bf16(g, device_sms(), s);
}

int ard_residual_backward(const float* x, const float* gout, const float* mean, const float* basis, const float* lam, float* dx, float* dlam,
                          long long rows, int D, int K, void* stream) {
    if (!x || !gout || !mean || !basis || !lam) return set_error(ARD_ERR_SHAPE, "ard_residual_backward: null argument");
    if (rows <= 0 || D <= 0 || K <= 0 || K > D || (D % 16) != 0) return set_error(ARD_ERR_SHAPE, "ard_residual_backward: bad shape rows=%lld D=%d K=%d", rows, D, K);
    cudaStream_t s = (cudaStream_t)stream;
    ResidualScratch& w = rscratch();
    const int Kp = (K + 15) & ~15;
    const int sms = device_sms();
    ARD_TRY(w.xb.ensure((size_t)rows * D * 2));
    ARD_TRY(w.gb.ensure((size_t)rows * D * 2));
    ARD_TRY(w.bb.ensure((size_t)Kp * D * 2));
    ARD_TRY(w.bbT.ensure((size_t)Kp * D * 2));
    ARD_TRY(w.c0.ensure((size_t)Kp * 4));
    ARD_TRY(w.lamp.ensure((size_t)Kp * 4));
    ARD_TRY(w.coef.ensure((size_t)rows * Kp * 4));
    ARD_TRY(w.gcoef.ensure((size_t)rows * Kp * 4));
    ARD_TRY(w.gsc.ensure((size_t)rows * Kp * 2));
    residual_pack_kernel<<<Kp, 256, 0, s>>>(basis, mean, lam, K, Kp, D, w.bb.as<__nv_bfloat16>(), w.bbT.as<__nv_bfloat16>(), w.c0.as<float>(),
                                            w.lamp.as<float>());
    ARD_CUDA(cudaGetLastError());
    count_launch();
    ARD_TRY(f32_to_bf16(gout, w.gb.as<__nv_bfloat16>(), rows * D, 1.0f, s));
    GemmArgs g;
    g.A = w.gb.as<__nv_bfloat16>(); g.lda = D; g.W = w.bb.as<__nv_bfloat16>(); g.ldw = D; g.out = w.gcoef.as<float>(); g.ldo = Kp;
    g.M = (int)rows; g.N = Kp; g.K = D;
    ARD_TRY(gemm_bf16(g, sms, s));                                   // gcoef = g B^T
    if (dlam != nullptr) {
        ARD_TRY(f32_to_bf16(x, w.xb.as<__nv_bfloat16>(), rows * D, 1.0f, s));
        g = GemmArgs();
        g.A = w.xb.as<__nv_bfloat16>(); g.lda = D; g.W = w.bb.as<__nv_bfloat16>(); g.ldw = D; g.out = w.coef.as<float>(); g.ldo = Kp;
        g.M = (int)rows; g.N = Kp; g.K = D; g.bias = w.c0.as<float>();
        ARD_TRY(gemm_bf16(g, sms, s));                               // coef = (x - mean) B^T
    }
    // dlam += colsum(coef * gcoef) (skipped when dlam is NULL: coef is then not read), gsc = bf16(gcoef * lam)
    ARD_TRY(lambda_grad(dlam ? w.coef.as<float>() : w.gcoef.as<float>(), w.gcoef.as<float>(), w.lamp.as<float>(), dlam, w.gsc.as<__nv_bfloat16>(),
                        rows, K, Kp, s));
    if (dx != nullptr) {
        g = GemmArgs();
        g.A = w.gsc.as<__nv_bfloat16>(); g.lda = Kp; g.W = w.bbT.as<__nv_bfloat16>(); g.ldw = Kp; g.out = dx; g.ldo = D;
        g.M = (int)rows; g.N = D; g.K = Kp;
        ARD_TRY(gemm_bf16(g, sms, s));                               // dx = (gcoef * lam) B
    }
    return 0;
}

int ard_head_forward(const float* emb, const float* W, const float* bias, int B, int N, int J, float* logits, void* stream) {
    if (!emb || !W || !logits || B <= 0 || N <= 0 || J <= 0) return set_error(ARD_ERR_SHAPE, "ard_head_forward: bad argument");
    return linear_small(emb, J, W, bias, logits, N, B, N, J, ARD_ACT_NONE, (cudaStream_t)stream);
}

int ard_ce_forward(const float* logits, const long long* labels, int B, int N, float* loss, float* dlogits, void* stream) {
    if (!logits || !labels || !loss || B <= 0 || N <= 0) return set_error(ARD_ERR_SHAPE, "ard_ce_forward: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    static DevBuf rows;
    ARD_TRY(rows.ensure((size_t)B * 4));
    ce_rows_kernel<<<(B + 7) / 8, 256, 0, s>>>(logits, labels, B, N, rows.as<float>(), dlogits);
    ARD_TRY(check_cuda(cudaGetLastError(), "ce_rows launch"));
    mean_rows_kernel<<<1, 256, 0, s>>>(rows.as<float>(), B, loss);
    return check_cuda(cudaGetLastError(), "ce_mean launch");
}

int ard_head_backward(const float* dlogits, const float* emb, const float* W, int B, int N, int J, float* d_emb, float* dW, float* db,
                      void* stream) {
    if (!dlogits || B <= 0 || N <= 0 || J <= 0) return set_error(ARD_ERR_SHAPE, "ard_head_backward: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (d_emb) {   // d_emb[b][j] = sum_n dz[b][n] W[n][j]
        if (!W) return set_error(ARD_ERR_SHAPE, "ard_head_backward: d_emb needs W");
        ARD_TRY(sgemm_strided(dlogits, 1, N, W, J, 1, d_emb, J, B, J, N, 1.0f, s));
    }
    if (dW) {      // dW[n][j] = sum_b dz[b][n] emb[b][j]
        if (!emb) return set_error(ARD_ERR_SHAPE, "ard_head_backward: dW needs emb");
        ARD_TRY(sgemm_strided(dlogits, N, 1, emb, J, 1, dW, J, N, J, B, 1.0f, s));
    }
    if (db) {
        colsum_kernel<<<(N + 127) / 128, 128, 0, s>>>(dlogits, B, N, db);
        ARD_TRY(check_cuda(cudaGetLastError(), "colsum launch"));
    }
    return 0;
}

int ard_eval_metrics(const float* scores, const long long* targets, long long n, int C, int k, long long* counts, long long* cm,
                     long long* preds, void* stream) {
    if (!scores || !targets || !counts || n < 0 || C <= 0 || k <= 0) return set_error(ARD_ERR_SHAPE, "ard_eval_metrics: bad argument");
    if (n == 0) return 0;
    eval_metrics_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(scores, targets, n, C, k, (unsigned long long*)counts,
                                                                                    (unsigned long long*)cm, preds);
    return check_cuda(cudaGetLastError(), "eval_metrics launch");
}

int ard_fill_clips(const void* flat, int src_is_pcm16, const long long* offsets, const int* lengths, int B, int max_len, int mode,
                   int quantize, float* out, void* stream) {
    if (!flat || !out || B <= 0 || max_len <= 0 || ((offsets == nullptr) != (lengths == nullptr)))
        return set_error(ARD_ERR_SHAPE, "ard_fill_clips: bad argument");
    if (mode < 0 || mode > 2) return set_error(ARD_ERR_NOTIMPL, "data_filling mode %d not implemented", mode);   // data.py:492-496
    dim3 grid(148, B);
    if (B > 65535) return set_error(ARD_ERR_SHAPE, "ard_fill_clips: at most 65535 clips per call");
    ProfScope ps(PROF_FRONTEND, (cudaStream_t)stream, 0.0, (double)B * max_len * (src_is_pcm16 ? 6.0 : 8.0));
    if (src_is_pcm16) fill_clips_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(flat, offsets, lengths, max_len, mode, quantize, out);
    else fill_clips_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(flat, offsets, lengths, max_len, mode, quantize, out);
    return check_cuda(cudaGetLastError(), "fill_clips launch");
}

}  // extern "C"
