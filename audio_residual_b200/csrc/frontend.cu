// Mel front end + patch embedding (HBM-bound, fp32 throughout).
//
//   stft_logmel_kernel : torchlibrosa Spectrogram + LogmelFilterBank as HTSAT uses them (htsat.py:681-687, :898-899):
//                        reflect-pad 512, periodic-Hann window, 1024-point DFT every 480 samples, |X|^2, mel projection
//                        (banded view of melW[513,64]), 10*log10(max(.,1e-10)). The reference runs the DFT as two
//                        Conv1d(1,513,k=1024) (2.1 GFLOP/clip) and writes the 513-bin spectrum; here two real frames share
//                        one complex 1024-point FFT held in a warp's registers (~0.05 GFLOP/clip) and only the 64 mel bins
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
constexpr int PAIRS_PER_CTA = 16;   // 8 warps x 2 frame pairs

// One warp per complex 1024-point FFT, data in registers (1024 = 32 x 32 Cooley-Tukey):
//   X[k1 + 32 k2] = sum_t W_1024^{t k1} W_32^{t k2} ( sum_j x[t + 32 j] W_32^{j k1} )
// lane t runs a 32-point FFT over j on x[t + 32 j], multiplies by W_1024^{t k1}, the warp transposes through a padded
// 32x33 shared tile (the only shared-memory traffic of the transform: 16 KB per FFT instead of 80 KB for five in-smem radix-4
// passes, which made the previous kernel shared-wavefront bound), lane k1 runs the second 32-point FFT over t.
// Two real frames ride in one complex transform (z = a + i b), split afterwards with Z[k] and Z[N-k] (one shuffle per bin).
__host__ __device__ constexpr int brev5(int x) { return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4); }
__device__ constexpr float C32[16] = {1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f, 7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f, 6.123233996e-17f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f, -7.071067812e-01f, -8.314696123e-01f, -9.238795325e-01f, -9.807852804e-01f};
__device__ constexpr float S32[16] = {0.000000000e+00f, 1.950903220e-01f, 3.826834324e-01f, 5.555702330e-01f, 7.071067812e-01f, 8.314696123e-01f, 9.238795325e-01f, 9.807852804e-01f, 1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f, 7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f};

// in-place radix-2 DIF, natural-order input, output X[k] lands in element brev5(k); W = exp(-2 pi i / 32)
ARD_DEVINL void fft32(float (&re)[32], float (&im)[32]) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
#pragma unroll
        for (int blk = 0; blk < 32; blk += 2 * half) {
#pragma unroll
            for (int k = 0; k < half; ++k) {
                const int i0 = blk + k, i1 = i0 + half;
                const int tw = k * (16 / half);
                const float ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
                re[i0] = ar + br;
                im[i0] = ai + bi;
                const float dr = ar - br, di = ai - bi;
                if (tw == 0) {
                    re[i1] = dr;
                    im[i1] = di;
                } else if (tw == 8) {            // * (-i)
                    re[i1] = di;
                    im[i1] = -dr;
                } else {                         // (dr + i di)(c - i s)
                    re[i1] = fmaf(dr, C32[tw], di * S32[tw]);
                    im[i1] = fmaf(di, C32[tw], -dr * S32[tw]);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) stft_logmel_kernel(const float* __restrict__ wave, int n_samples, int frames,
                                                         const float* __restrict__ window, const float2* __restrict__ twiddle,
                                                         const float* __restrict__ melw, const int* __restrict__ mstart,
                                                         const int* __restrict__ mlen, int band_max,
                                                         const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                                         float* __restrict__ out, long long out_clip_stride, int replicate, int quantize) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float2 tw2[32 * 32];        // tw2[k1][t] = W_1024^{t k1}
    __shared__ float win[NFFT];
    __shared__ float wbuf[8][32 * 33];     // per-warp transpose tile, later the two power spectra
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long b = blockIdx.y;
    const float* x = wave + b * n_samples;
    for (int i = tid; i < NFFT; i += 256) {
        tw2[i] = __ldg(twiddle + (((i & 31) * (i >> 5)) & (NFFT - 1)));
        win[i] = __ldg(window + i);
    }
    __syncthreads();
    float* buf = wbuf[warp];
    const int npairs = (frames + 1) >> 1;
    const int p_end = min((int)(blockIdx.x + 1) * PAIRS_PER_CTA, npairs);
    for (int pr = blockIdx.x * PAIRS_PER_CTA + warp; pr < p_end; pr += 8) {
        const int fa = 2 * pr, fb = fa + 1;
        const int base_a = fa * HOPS - NFFT / 2;
        float re[32], im[32];
        // ---- samples: lane t holds n = t + 32 j. Frame b is frame a advanced by 480 = 15 * 32 samples: its j < 17 values are
        // frame a's j + 15 values of the same lane, only j >= 17 is loaded. Reflect padding (F.pad mode='reflect') at clip ends.
        const bool interior = base_a >= 0 && base_a + HOPS + NFFT <= n_samples;   // warp-uniform
        if (interior) {
#pragma unroll
            for (int j = 0; j < 32; ++j) re[j] = __ldg(x + base_a + lane + 32 * j);
#pragma unroll
            for (int j = 0; j < 17; ++j) im[j] = re[j + 15];
#pragma unroll
            for (int j = 17; j < 32; ++j) im[j] = __ldg(x + base_a + HOPS + lane + 32 * j);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                int idx = base_a + lane + 32 * j;
                idx = idx < 0 ? -idx : (idx >= n_samples ? 2 * (n_samples - 1) - idx : idx);
                re[j] = __ldg(x + idx);
                idx = base_a + HOPS + lane + 32 * j;
                idx = idx < 0 ? -idx : (idx >= n_samples ? 2 * (n_samples - 1) - idx : idx);
                im[j] = __ldg(x + idx);
            }
        }
        if (fb >= frames) {
#pragma unroll
            for (int j = 0; j < 32; ++j) im[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float a0 = re[j], b0 = im[j];
            if (quantize) {   // quantize_tensor, src/residual.py:210-212
                a0 = truncf(fminf(fmaxf(a0, -1.f), 1.f) * 32767.0f) / 32767.0f;
                b0 = truncf(fminf(fmaxf(b0, -1.f), 1.f) * 32767.0f) / 32767.0f;
            }
            const float w = win[lane + 32 * j];
            re[j] = a0 * w;
            im[j] = b0 * w;
        }
        fft32(re, im);                                   // element r = Y_t[k1 = brev5(r)]
#pragma unroll
        for (int r = 0; r < 32; ++r) {                   // * W_1024^{t k1}
            const float2 w = tw2[brev5(r) * 32 + lane];
            const float yr = re[r] * w.x - im[r] * w.y;
            im[r] = fmaf(re[r], w.y, im[r] * w.x);
            re[r] = yr;
        }
        __syncwarp();                                    // previous pair's mel stage is done with buf
#pragma unroll
        for (int r = 0; r < 32; ++r) buf[brev5(r) * 33 + lane] = re[r];
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 32; ++t) re[t] = buf[lane * 33 + t];
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; ++r) buf[brev5(r) * 33 + lane] = im[r];
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 32; ++t) im[t] = buf[lane * 33 + t];
        __syncwarp();
        fft32(re, im);                                   // lane = k1, element brev5(k2) = Z[k1 + 32 k2]
        // ---- split the two real spectra and take powers: Xa = (Z[k] + conj Z[N-k]) / 2, Xb = (Z[k] - conj Z[N-k]) / (2i),
        // N - k = ((32 - k1) & 31) + 32 (31 - k2) for k1 > 0, and 32 ((32 - k2) & 31) for k1 = 0. Bins 0..512 only.
        const int plane = (32 - lane) & 31;
#pragma unroll
        for (int k2 = 0; k2 <= 16; ++k2) {
            const float zr = re[brev5(k2)], zi = im[brev5(k2)];
            float qr = __shfl_sync(0xffffffffu, re[brev5(31 - k2)], plane);
            float qi = __shfl_sync(0xffffffffu, im[brev5(31 - k2)], plane);
            if (lane == 0) {
                qr = re[brev5((32 - k2) & 31)];
                qi = im[brev5((32 - k2) & 31)];
            }
            const float sr = zr + qr, di = zi - qi, si = zi + qi, dr = zr - qr;
            const int k = lane + 32 * k2;
            if (k <= NFFT / 2) {
                buf[k] = 0.25f * fmaf(sr, sr, di * di);
                buf[520 + k] = 0.25f * fmaf(si, si, dr * dr);
            }
        }
        __syncwarp();
        // ---- banded mel projection + log: lane handles mel bins lane and lane + 32 of both frames
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int m = lane + 32 * half;
            const int st = __ldg(mstart + m), ln = __ldg(mlen + m);
            const float* wrow = melw + m * band_max;
            float acc0 = 0.f, acc1 = 0.f;
            for (int q = 0; q < ln; ++q) {
                const float w = __ldg(wrow + q);
                acc0 = fmaf(buf[st + q], w, acc0);
                acc1 = fmaf(buf[520 + st + q], w, acc1);
            }
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                const int frame = fa + f;
                if (frame < frames) {
                    float v = 10.0f * log10f(fmaxf(f ? acc1 : acc0, 1e-10f));   // ref=1.0 -> "- 10*log10(max(amin, ref))" is exactly 0
                    if (bn_scale != nullptr) v = fmaf(v, bn_scale[m], bn_shift[m]);
                    float* o = out + b * out_clip_stride + (long long)frame * NMEL + m;
                    for (int rpl = 0; rpl < replicate; ++rpl) o[(long long)rpl * frames * NMEL] = v;
                }
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
    ARD_CUDA(enqueue_pdl(stft_logmel_kernel, grid, dim3(256), 0, s, wave, n_samples, frames, window, twiddle, mel.w, mel.start, mel.len, mel.band_max,
                        bn_scale, bn_shift, out, out_clip_stride, replicate, quantize));
    return check_cuda(cudaGetLastError(), "stft_logmel launch");
}

// ------------------------------------------------------------------------------------------------ patch embed
// upsample_bicubic2d coefficients (A = -0.75), as in ATen's cubic_convolution1/2
ARD_DEVINL float cc1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
ARD_DEVINL float cc2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

constexpr int PE_TOK_PER_WARP = 32;   // consecutive tokens (same patch row) handled by one warp

template <int CPL>   // channels per lane: C = 32 * CPL
__global__ void __launch_bounds__(256) patch_embed_ln_kernel(const float* __restrict__ mel, long long clip_stride, int frames,
                                                            const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                                            const float* __restrict__ wconv, const float* __restrict__ bconv,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ out, long long ntokens) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int C = 32 * CPL;
    __shared__ __align__(16) float pixbuf[8][2][32];   // per warp, double-buffered (the next step's store must not race this step's reads)
    const int lane = threadIdx.x & 31;
    // each lane keeps the 4x4 conv weights, bias and LayerNorm affine of its CPL channels in registers for the whole run of
    // tokens (16 * CPL + 3 * CPL values): no shared memory, no per-block weight staging
    float wr[16][CPL], bq[CPL], gq[CPL], eq[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const int c = lane + 32 * q;
        bq[q] = __ldg(bconv + c);
        gq[q] = __ldg(gamma + c);
        eq[q] = __ldg(beta + c);
#pragma unroll
        for (int n = 0; n < 16; ++n) wr[n][q] = __ldg(wconv + c * 16 + n);
    }
    const long long tok0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * PE_TOK_PER_WARP;
    if (tok0 >= ntokens) return;
    // The warp's 32 tokens share one patch row (64 tokens per row, runs are 32-aligned): clip, quarter r, mel bins f are fixed.
    // Two tokens per step: lanes 0..15 produce the 4x4 pixels of the even token, lanes 16..31 those of the odd one
    // (image row 4ph+i -> (quarter r, mel bin f), col 4pw+j -> time). The next step's mel taps are loaded before this step's
    // arithmetic so the L2 latency overlaps it.
    const long long b = tok0 >> 12;
    const int t0 = (int)(tok0 & 4095);
    const int ph = t0 >> 6, pw0 = t0 & 63;
    const int i = (lane >> 2) & 3, j = lane & 3, sub = lane >> 4;
    const int r = ph >> 4;
    const int f = ((ph & 15) << 2) + i;
    const float scale = (float)(frames - 1) / (float)(1024 - 1);   // align_corners=True
    const float A = -0.75f;
    const float* mp = mel + b * clip_stride + f;
    const float sc = bn_scale ? __ldg(bn_scale + f) : 1.f, sh = bn_shift ? __ldg(bn_shift + f) : 0.f;
    auto taps = [&](int it, float (&v)[4]) {
        const int tau = r * 256 + (pw0 + it + sub) * 4 + j;        // index on the 1024-frame (interpolated) time axis
        const int x0 = (int)floorf(scale * (float)tau);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int xi = x0 - 1 + k;
            xi = xi < 0 ? 0 : (xi > frames - 1 ? frames - 1 : xi);
            v[k] = __ldg(mp + (long long)xi * NMEL);
        }
    };
    float v[4], vn[4];
    taps(0, v);
    for (int it = 0; it < PE_TOK_PER_WARP; it += 2) {
        if (it + 2 < PE_TOK_PER_WARP) taps(it + 2, vn);
        float pix;
        {
            const int tau = r * 256 + (pw0 + it + sub) * 4 + j;
            const float real = scale * (float)tau;
            const float tt = real - floorf(real);
            pix = fmaf(v[0], sc, sh) * cc2(tt + 1.f, A);                   // bn0 before the interpolation (htsat.py:900-902)
            pix = fmaf(fmaf(v[1], sc, sh), cc1(tt, A), pix);
            pix = fmaf(fmaf(v[2], sc, sh), cc1(1.f - tt, A), pix);
            pix = fmaf(fmaf(v[3], sc, sh), cc2(2.f - tt, A), pix);
        }
        float acc[2][CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) acc[0][q] = acc[1][q] = bq[q];
        // broadcast the 2 x 16 pixels through shared memory: one store + eight broadcast LDS.128 instead of 32 shuffles (the
        // kernel was bound by the shuffle pipe: 26 SHFL per token at one warp-shuffle per clock per SM)
        float* pb = pixbuf[threadIdx.x >> 5][(it >> 1) & 1];
        pb[lane] = pix;
        __syncwarp();
#pragma unroll
        for (int n4 = 0; n4 < 4; ++n4) {
            const float4 a4 = *reinterpret_cast<const float4*>(pb + n4 * 4);
            const float4 b4 = *reinterpret_cast<const float4*>(pb + 16 + n4 * 4);
            const float pa[4] = {a4.x, a4.y, a4.z, a4.w}, pc[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    acc[0][q] = fmaf(pa[k], wr[n4 * 4 + k][q], acc[0][q]);
                    acc[1][q] = fmaf(pc[k], wr[n4 * 4 + k][q], acc[1][q]);
                }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long tok = tok0 + it + u;
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < CPL; ++q) s += acc[u][q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s * (1.0f / C);
            float var = 0.f;
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const float d = acc[u][q] - mean;
                var = fmaf(d, d, var);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
            const float rstd = rsqrtf(var * (1.0f / C) + 1e-5f);
#pragma unroll
            for (int q = 0; q < CPL; ++q) out[tok * C + lane + 32 * q] = fmaf((acc[u][q] - mean) * rstd, gq[q], eq[q]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = vn[k];
    }
}

int patch_embed_ln(const float* logmel, long long clip_stride, int frames, const float* bn_scale, const float* bn_shift, const float* w,
                   const float* bias, const float* gamma, const float* beta, float* out, int B, int C, cudaStream_t s) {
    if (B <= 0) return 0;
    if (frames > 1024) return set_error(ARD_ERR_SHAPE, "the wav size should less than or equal to the swin input size");  // htsat.py:852
    const long long ntok = (long long)B * 4096;
    const unsigned grid = (unsigned)((ntok + 8 * PE_TOK_PER_WARP - 1) / (8 * PE_TOK_PER_WARP));
    ProfScope ps(PROF_FRONTEND, s, (double)ntok * (2.0 * 16 * C + 16 * 8 + 8.0 * C), 4.0 * B * frames * 64 + 4.0 * ntok * C);
    if (C == 96)
        ARD_CUDA(enqueue_pdl(patch_embed_ln_kernel<3>, dim3(grid), dim3(256), 0, s, logmel, clip_stride, frames, bn_scale, bn_shift, w, bias, gamma, beta, out, ntok));
    else if (C == 128)
        ARD_CUDA(enqueue_pdl(patch_embed_ln_kernel<4>, dim3(grid), dim3(256), 0, s, logmel, clip_stride, frames, bn_scale, bn_shift, w, bias, gamma, beta, out, ntok));
    else
        return set_error(ARD_ERR_SHAPE, "patch_embed: unsupported embed_dim %d", C);
    return check_cuda(cudaGetLastError(), "patch_embed_ln launch");
}

}  // namespace ard
