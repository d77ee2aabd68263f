// Mel front end + patch embedding (HBM-bound, fp32 throughout).
//
//   stft_logmel_kernel : torchlibrosa Spectrogram + LogmelFilterBank as HTSAT uses them (htsat.py:681-687, :898-899):
//                        reflect-pad 512, periodic-Hann window, 1024-point DFT every 480 samples, |X|^2, mel projection
//                        (banded view of melW[513,64]), 10*log10(max(.,1e-10)). The reference runs the DFT as two
//                        Conv1d(1,513,k=1024) (2.1 GFLOP/clip) and writes the 513-bin spectrum; here two real frames share
//                        one complex radix-4 Stockham FFT in shared memory (~0.05 GFLOP/clip) and only the 64 mel bins
//                        reach HBM. Also serves the fusion featuriser's get_mel (data.py:363-399; same formula, htk filters).
//   patch_embed_ln_kernel : bn0 (eval) + reshape_wav2img (bicubic 1001->1024 along time, fold into 4 frequency-stacked
//                        quarters; htsat.py:848-863, :900-902) + PatchEmbed conv 4x4/4 + LayerNorm (htsat.py:136-143) fused:
//                        the 256x256 image never exists in memory.
#include "ard_common.cuh"
#include "ard_internal.h"

namespace ard {

constexpr int NFFT = 1024;
constexpr int HOPS = 480;
constexpr int NBINS = 513;
constexpr int NMEL = 64;
constexpr int PAIRS_PER_CTA = 8;

ARD_DEVINL float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// FFT buffers are indexed through PADI: one float2 of padding after every 4 elements. The radix-4 Stockham passes store with
// element strides of 4/16/64: unpadded that is an 8-way bank conflict on the first passes (ncu: 56% of all shared wavefronts
// of this kernel were conflict replays), padded the stride-4 pattern is conflict-free and the others at most 2-way.
#define PADI(i) ((i) + ((i) >> 2))
constexpr int NFFT_PAD = NFFT + NFFT / 4;

__global__ void __launch_bounds__(256) stft_logmel_kernel(const float* __restrict__ wave, int n_samples, int frames,
                                                         const float* __restrict__ window, const float2* __restrict__ twiddle,
                                                         const float* __restrict__ melw, const int* __restrict__ mstart,
                                                         const int* __restrict__ mlen, int band_max,
                                                         const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                                         float* __restrict__ out, long long out_clip_stride, int replicate, int quantize) {
    __shared__ float2 bufA[NFFT_PAD];
    __shared__ float2 bufB[NFFT_PAD];
    __shared__ float2 tw[NFFT];
    __shared__ float win[NFFT];
    __shared__ float pw[2][NBINS + 3];
    const int tid = threadIdx.x;
    const long long b = blockIdx.y;
    const float* x = wave + b * n_samples;
    for (int i = tid; i < NFFT; i += 256) {
        tw[i] = twiddle[i];
        win[i] = window[i];
    }
    const int npairs = (frames + 1) >> 1;
    const int p_begin = blockIdx.x * PAIRS_PER_CTA;
    const int p_end = min(p_begin + PAIRS_PER_CTA, npairs);
    // samples of a frame pair: reflect padding (F.pad mode='reflect', n_fft/2 each side) only matters near the clip ends
    float va[4], vb[4];
    auto load_pair = [&](int pr) {
        const int fa = 2 * pr, fb = 2 * pr + 1;
        const int base_a = fa * HOPS - NFFT / 2;
        const bool interior = base_a >= 0 && base_a + HOPS + NFFT <= n_samples && fb < frames;   // block-uniform
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int n = tid + 256 * r;
            vb[r] = 0.f;
            if (interior) {
                va[r] = __ldg(x + base_a + n);
                vb[r] = __ldg(x + base_a + HOPS + n);
            } else {
                int idx = base_a + n;
                idx = idx < 0 ? -idx : (idx >= n_samples ? 2 * (n_samples - 1) - idx : idx);
                va[r] = __ldg(x + idx);
                if (fb < frames) {
                    idx = base_a + HOPS + n;
                    idx = idx < 0 ? -idx : (idx >= n_samples ? 2 * (n_samples - 1) - idx : idx);
                    vb[r] = __ldg(x + idx);
                }
            }
        }
    };
    if (p_begin < p_end) load_pair(p_begin);
    for (int pr = p_begin; pr < p_end; ++pr) {
        const int fa = 2 * pr;
        __syncthreads();
        // z[n] = w[n] * (xa[n] + i xb[n])
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int n = tid + 256 * r;
            float a0 = va[r], b0 = vb[r];
            if (quantize) {   // quantize_tensor, src/residual.py:210-212
                a0 = truncf(fminf(fmaxf(a0, -1.f), 1.f) * 32767.0f) / 32767.0f;
                b0 = truncf(fminf(fmaxf(b0, -1.f), 1.f) * 32767.0f) / 32767.0f;
            }
            const float w = win[n];
            bufA[PADI(n)] = make_float2(a0 * w, b0 * w);
        }
        if (pr + 1 < p_end) load_pair(pr + 1);       // next pair's samples are in flight during this pair's FFT
        __syncthreads();
        // radix-4 Stockham autosort FFT, 5 passes (p = 1,4,16,64,256), natural-order output
        float2* src = bufA;
        float2* dst = bufB;
#pragma unroll
        for (int pass = 0; pass < 5; ++pass) {
            const int p = 1 << (2 * pass);
            const int k = tid & (p - 1);
            const int j = ((tid - k) << 2) + k;
            const int tstep = (NFFT / 4) / p * k;   // twiddle index for exp(-2 pi i k / (4p))
            float2 u0 = src[PADI(tid)], u1 = src[PADI(tid + 256)], u2 = src[PADI(tid + 512)], u3 = src[PADI(tid + 768)];
            if (pass > 0) {
                u1 = cmul(u1, tw[tstep]);
                u2 = cmul(u2, tw[2 * tstep]);
                u3 = cmul(u3, tw[3 * tstep]);
            }
            const float2 v0 = make_float2(u0.x + u2.x, u0.y + u2.y);
            const float2 v1 = make_float2(u0.x - u2.x, u0.y - u2.y);
            const float2 v2 = make_float2(u1.x + u3.x, u1.y + u3.y);
            const float2 d = make_float2(u1.x - u3.x, u1.y - u3.y);
            const float2 v3 = make_float2(d.y, -d.x);   // (u1 - u3) * (-i)
            dst[PADI(j)] = make_float2(v0.x + v2.x, v0.y + v2.y);
            dst[PADI(j + p)] = make_float2(v1.x + v3.x, v1.y + v3.y);
            dst[PADI(j + 2 * p)] = make_float2(v0.x - v2.x, v0.y - v2.y);
            dst[PADI(j + 3 * p)] = make_float2(v1.x - v3.x, v1.y - v3.y);
            __syncthreads();
            float2* t = src; src = dst; dst = t;
        }
        // split the two real spectra and take powers: Xa = (Z[k] + conj Z[N-k]) / 2, Xb = (Z[k] - conj Z[N-k]) / (2i)
        for (int k = tid; k < NBINS; k += 256) {
            const float2 z = src[PADI(k)];
            const int kc = (NFFT - k) & (NFFT - 1);
            const float2 zc = src[PADI(kc)];
            const float ar = 0.5f * (z.x + zc.x), ai = 0.5f * (z.y - zc.y);
            const float br = 0.5f * (z.y + zc.y), bi = 0.5f * (zc.x - z.x);
            pw[0][k] = ar * ar + ai * ai;
            pw[1][k] = br * br + bi * bi;
        }
        __syncthreads();
        if (tid < 2 * NMEL) {
            const int f = tid >> 6, m = tid & 63;
            const int frame = fa + f;
            if (frame < frames) {
                const int st = mstart[m], ln = mlen[m];
                const float* wrow = melw + m * band_max;
                float acc = 0.f;
                for (int q = 0; q < ln; ++q) acc = fmaf(pw[f][st + q], __ldg(wrow + q), acc);
                float v = 10.0f * log10f(fmaxf(acc, 1e-10f));   // ref=1.0 -> "- 10*log10(max(amin, ref))" is exactly 0
                if (bn_scale != nullptr) v = fmaf(v, bn_scale[m], bn_shift[m]);
                float* o = out + b * out_clip_stride + (long long)frame * NMEL + m;
                for (int rpl = 0; rpl < replicate; ++rpl) o[(long long)rpl * frames * NMEL] = v;
            }
        }
    }
}

int stft_logmel(const float* wave, int B, int n_samples, const float* window, const float2* twiddle, const MelBands& mel,
                const float* bn_scale, const float* bn_shift, float* out, long long out_clip_stride, int replicate, int quantize,
                cudaStream_t s) {
    if (B <= 0) return 0;
    if (n_samples <= NFFT / 2) return set_error(ARD_ERR_SHAPE, "stft: clip too short for reflect padding (%d samples)", n_samples);
    const int frames = n_samples / HOPS + 1;
    const int npairs = (frames + 1) / 2;
    if (out_clip_stride <= 0) out_clip_stride = (long long)frames * NMEL;
    if (replicate < 1) replicate = 1;
    dim3 grid((npairs + PAIRS_PER_CTA - 1) / PAIRS_PER_CTA, B);
    ProfScope ps(PROF_FRONTEND, s, (double)B * npairs * (5.0 * 1024 * 10 + 2.0 * 2 * 1100), 4.0 * B * n_samples + 4.0 * B * frames * 64 * replicate);
    stft_logmel_kernel<<<grid, 256, 0, s>>>(wave, n_samples, frames, window, twiddle, mel.w, mel.start, mel.len, mel.band_max, bn_scale,
                                           bn_shift, out, out_clip_stride, replicate, quantize);
    return check_cuda(cudaGetLastError(), "stft_logmel launch");
}

// ------------------------------------------------------------------------------------------------ patch embed
// upsample_bicubic2d coefficients (A = -0.75), as in ATen's cubic_convolution1/2
ARD_DEVINL float cc1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
ARD_DEVINL float cc2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

template <int CPL>   // channels per lane: C = 32 * CPL
__global__ void __launch_bounds__(256) patch_embed_ln_kernel(const float* __restrict__ mel, long long clip_stride, int frames,
                                                            const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                                            const float* __restrict__ wconv, const float* __restrict__ bconv,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ out, long long ntokens) {
    constexpr int C = 32 * CPL;
    __shared__ float wt[16][C];   // transposed conv weight: wt[kh*4+kw][c]
    for (int i = threadIdx.x; i < 16 * C; i += blockDim.x) {
        const int c = i / 16, n = i % 16;
        wt[n][c] = wconv[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tok >= ntokens) return;
    const long long b = tok >> 12;
    const int t = (int)(tok & 4095);
    const int ph = t >> 6, pwi = t & 63;
    // lanes 0..15 each produce one pixel of the 4x4 patch: image row 4ph+i -> (quarter r, mel bin f), col 4pw+j -> time
    float pix = 0.f;
    {
        const int i = (lane >> 2) & 3, j = lane & 3;
        const int r = ph >> 4;
        const int f = ((ph & 15) << 2) + i;
        const int tau = r * 256 + pwi * 4 + j;                 // index on the 1024-frame (interpolated) time axis
        const float scale = (float)(frames - 1) / (float)(1024 - 1);   // align_corners=True
        const float real = scale * (float)tau;
        const int x0 = (int)floorf(real);
        const float tt = real - (float)x0;
        const float A = -0.75f;
        const float cw[4] = {cc2(tt + 1.f, A), cc1(tt, A), cc1(1.f - tt, A), cc2(2.f - tt, A)};
        const float* mp = mel + b * clip_stride + f;
        const float sc = bn_scale ? bn_scale[f] : 1.f, sh = bn_shift ? bn_shift[f] : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int xi = x0 - 1 + k;
            xi = xi < 0 ? 0 : (xi > frames - 1 ? frames - 1 : xi);
            const float v = fmaf(__ldg(mp + (long long)xi * NMEL), sc, sh);     // bn0 before the interpolation (htsat.py:900-902)
            pix = fmaf(v, cw[k], pix);
        }
    }
    float acc[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) acc[q] = __ldg(bconv + lane + 32 * q);
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        const float pv = __shfl_sync(0xffffffffu, pix, n);
#pragma unroll
        for (int q = 0; q < CPL; ++q) acc[q] = fmaf(pv, wt[n][lane + 32 * q], acc[q]);
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < CPL; ++q) s += acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / C);
    float var = 0.f;
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const float d = acc[q] - mean;
        var = fmaf(d, d, var);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var * (1.0f / C) + 1e-5f);
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const int c = lane + 32 * q;
        out[tok * C + c] = fmaf((acc[q] - mean) * rstd, __ldg(gamma + c), __ldg(beta + c));
    }
}

int patch_embed_ln(const float* logmel, long long clip_stride, int frames, const float* bn_scale, const float* bn_shift, const float* w,
                   const float* bias, const float* gamma, const float* beta, float* out, int B, int C, cudaStream_t s) {
    if (B <= 0) return 0;
    if (frames > 1024) return set_error(ARD_ERR_SHAPE, "the wav size should less than or equal to the swin input size");  // htsat.py:852
    const long long ntok = (long long)B * 4096;
    const unsigned grid = (unsigned)((ntok + 7) / 8);
    ProfScope ps(PROF_FRONTEND, s, (double)ntok * (2.0 * 16 * C + 16 * 8 + 8.0 * C), 4.0 * B * frames * 64 + 4.0 * ntok * C);
    if (C == 96)
        patch_embed_ln_kernel<3><<<grid, 256, 0, s>>>(logmel, clip_stride, frames, bn_scale, bn_shift, w, bias, gamma, beta, out, ntok);
    else if (C == 128)
        patch_embed_ln_kernel<4><<<grid, 256, 0, s>>>(logmel, clip_stride, frames, bn_scale, bn_shift, w, bias, gamma, beta, out, ntok);
    else
        return set_error(ARD_ERR_SHAPE, "patch_embed: unsupported embed_dim %d", C);
    return check_cuda(cudaGetLastError(), "patch_embed_ln launch");
}

}  // namespace ard
