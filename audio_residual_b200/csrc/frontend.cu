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
#include <cstdlib>

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

// Tensor-core form of the same op. The 4x4 conv is a [tokens, 16] x [16, C] product: exactly one k-step of mma.sync
// m16n8k16. To keep fp32-grade accuracy on bf16 tensor cores both operands are split, x = hi + lo with hi = bf16(x),
// lo = bf16(x - hi), and three products are accumulated (hi*hi + lo*hi + hi*lo: relative error ~2^-16). A warp walks runs of 32
// consecutive tokens of one patch row: the 32 lanes compute the 2 x 16 pixels of two tokens per step as before (bn0 ->
// bicubic -> fold), park them as split bf16 in a padded shared tile, then run two 16-token m-tiles: A fragments by
// ldmatrix, the weight fragments live in registers for the warp's whole life, LayerNorm on the accumulators (a row's 96/128
// channels sit in the 4 lanes of a quad: two shuffles per reduction instead of five per token), 8-byte stores.
// The SIMT kernel above spent 26 shuffles and ~140 instructions per token (360 us at B = 256); this one ~45.
constexpr int PE_RUNS_PER_WARP = 4;     // 128 tokens per warp: amortises the weight-fragment set-up
constexpr int PE_APITCH = 48;           // bytes per token row of the pixel tile (16 bf16 + pad: conflict-free ldmatrix)

template <int NT>   // n-tiles of 8 channels: C = 8 * NT (12 or 16)
__global__ void __launch_bounds__(256) patch_embed_ln_mma_kernel(const float* __restrict__ mel, long long clip_stride, int frames,
                                                                const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                                                const float* __restrict__ wconv, const float* __restrict__ bconv,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float* __restrict__ out, long long ntokens) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int C = 8 * NT;
    __shared__ __align__(16) uint8_t atile[8][2][32 * PE_APITCH];   // per warp: hi / lo pixel tiles [32 tokens][16 pixels] bf16
    __shared__ float2 aff[3][C / 2];                                // conv bias, LayerNorm gamma, beta as channel pairs
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < C / 2; i += blockDim.x) {
        aff[0][i] = make_float2(__ldg(bconv + 2 * i), __ldg(bconv + 2 * i + 1));
        aff[1][i] = make_float2(__ldg(gamma + 2 * i), __ldg(gamma + 2 * i + 1));
        aff[2][i] = make_float2(__ldg(beta + 2 * i), __ldg(beta + 2 * i + 1));
    }
    // B fragments of W^T [k = pixel][n = channel] (m16n8k16 col-major B): lane holds k = (lane%4)*2 + {0,1} (+8), n = lane/4
    uint32_t bh[NT][2], bl[NT][2];
    {
        const int k0 = (lane & 3) * 2, n = lane >> 2;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const float* wp = wconv + (nt * 8 + n) * 16;
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
                const float w0 = __ldg(wp + k0 + 8 * hlf), w1 = __ldg(wp + k0 + 8 * hlf + 1);
                const __nv_bfloat16 h0 = __float2bfloat16_rn(w0), h1 = __float2bfloat16_rn(w1);
                bh[nt][hlf] = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
                bl[nt][hlf] = pack_bf16x2(w0 - __bfloat162float(h0), w1 - __bfloat162float(h1));
            }
        }
    }
    __syncthreads();
    uint8_t* at_hi = atile[warp][0];
    uint8_t* at_lo = atile[warp][1];
    const int i = (lane >> 2) & 3, j = lane & 3, sub = lane >> 4;
    const float scale = (float)(frames - 1) / (float)(1024 - 1);   // align_corners=True
    const float A = -0.75f;
    const long long warp_tok0 = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * (32 * PE_RUNS_PER_WARP);
#pragma unroll 1
    for (int run = 0; run < PE_RUNS_PER_WARP; ++run) {
        const long long tok0 = warp_tok0 + run * 32;
        if (tok0 >= ntokens) break;
        // The run's 32 tokens share one patch row (64 tokens per row, runs are 32-aligned): clip, quarter r, mel bins f are fixed.
        // (Staging the run's ~128 source frames in shared memory first was measured slower: 0.97 vs 0.79 ms for the front end -
        // the staging loads are a serial latency per run, while these taps are prefetched one step ahead.)
        const long long b = tok0 >> 12;
        const int t0 = (int)(tok0 & 4095);
        const int ph = t0 >> 6, pw0 = t0 & 63;
        const int r = ph >> 4;
        const int f = ((ph & 15) << 2) + i;
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
        __syncwarp();                                  // the previous run's ldmatrix reads of the pixel tiles are done
#pragma unroll 1
        for (int it = 0; it < 32; it += 2) {
            if (it + 2 < 32) taps(it + 2, vn);
            const int tau = r * 256 + (pw0 + it + sub) * 4 + j;
            const float real = scale * (float)tau;
            const float tt = real - floorf(real);
            float pix = fmaf(v[0], sc, sh) * cc2(tt + 1.f, A);             // bn0 before the interpolation (htsat.py:900-902)
            pix = fmaf(fmaf(v[1], sc, sh), cc1(tt, A), pix);
            pix = fmaf(fmaf(v[2], sc, sh), cc1(1.f - tt, A), pix);
            pix = fmaf(fmaf(v[3], sc, sh), cc2(2.f - tt, A), pix);
            const __nv_bfloat16 hi = __float2bfloat16_rn(pix);
            const __nv_bfloat16 lo = __float2bfloat16_rn(pix - __bfloat162float(hi));
            const int off = (it + sub) * PE_APITCH + (lane & 15) * 2;      // token it+sub, pixel n = 4 i + j
            *reinterpret_cast<__nv_bfloat16*>(at_hi + off) = hi;
            *reinterpret_cast<__nv_bfloat16*>(at_lo + off) = lo;
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = vn[k];
        }
        __syncwarp();
#pragma unroll 1
        for (int mt = 0; mt < 2; ++mt) {
            // A fragments (16 tokens x 16 pixels): ldmatrix.x4, lane -> row (lane & 15), 16-byte half (lane >> 4)
            uint32_t ah[4], al[4];
            const uint32_t aoff = (uint32_t)((mt * 16 + (lane & 15)) * PE_APITCH + (lane >> 4) * 16);
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(ah[0]), "=r"(ah[1]), "=r"(ah[2]), "=r"(ah[3]) : "r"(smem_u32(at_hi) + aoff));
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(al[0]), "=r"(al[1]), "=r"(al[2]), "=r"(al[3]) : "r"(smem_u32(at_lo) + aoff));
            float acc[NT][4];
            const int cp = lane & 3;                   // this lane's channel pair inside an n-tile: channels 8 nt + 2 cp, +1
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 bb = aff[0][nt * 4 + cp];
                acc[nt][0] = bb.x; acc[nt][1] = bb.y; acc[nt][2] = bb.x; acc[nt][3] = bb.y;
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                             : "r"(ah[0]), "r"(ah[1]), "r"(ah[2]), "r"(ah[3]), "r"(bh[nt][0]), "r"(bh[nt][1]));
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                             : "r"(al[0]), "r"(al[1]), "r"(al[2]), "r"(al[3]), "r"(bh[nt][0]), "r"(bh[nt][1]));
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                             : "r"(ah[0]), "r"(ah[1]), "r"(ah[2]), "r"(ah[3]), "r"(bl[nt][0]), "r"(bl[nt][1]));
            }
            // LayerNorm over the C channels of rows g = lane / 4 (acc[.][0..1]) and g + 8 (acc[.][2..3]): the row lives in the quad
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) { s0 += acc[nt][0] + acc[nt][1]; s1 += acc[nt][2] + acc[nt][3]; }
            s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
            const float m0 = s0 * (1.0f / C), m1 = s1 * (1.0f / C);
            float q0 = 0.f, q1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                acc[nt][0] -= m0; acc[nt][1] -= m0; acc[nt][2] -= m1; acc[nt][3] -= m1;
                q0 = fmaf(acc[nt][0], acc[nt][0], fmaf(acc[nt][1], acc[nt][1], q0));
                q1 = fmaf(acc[nt][2], acc[nt][2], fmaf(acc[nt][3], acc[nt][3], q1));
            }
            q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
            q0 += __shfl_xor_sync(0xffffffffu, q0, 2); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
            const float r0 = rsqrtf(q0 * (1.0f / C) + 1e-5f), r1 = rsqrtf(q1 * (1.0f / C) + 1e-5f);
            float* o0 = out + (tok0 + mt * 16 + (lane >> 2)) * C + cp * 2;
            float* o1 = o0 + 8 * C;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 gg = aff[1][nt * 4 + cp], be = aff[2][nt * 4 + cp];
                *reinterpret_cast<float2*>(o0 + nt * 8) = make_float2(fmaf(acc[nt][0] * r0, gg.x, be.x), fmaf(acc[nt][1] * r0, gg.y, be.y));
                *reinterpret_cast<float2*>(o1 + nt * 8) = make_float2(fmaf(acc[nt][2] * r1, gg.x, be.x), fmaf(acc[nt][3] * r1, gg.y, be.y));
            }
        }
    }
}

int patch_embed_ln(const float* logmel, long long clip_stride, int frames, const float* bn_scale, const float* bn_shift, const float* w,
                   const float* bias, const float* gamma, const float* beta, float* out, int B, int C, cudaStream_t s) {
    if (B <= 0) return 0;
    if (frames > 1024) return set_error(ARD_ERR_SHAPE, "the wav size should less than or equal to the swin input size");  // htsat.py:852
    const long long ntok = (long long)B * 4096;
    const unsigned grid = (unsigned)((ntok + 8 * PE_TOK_PER_WARP - 1) / (8 * PE_TOK_PER_WARP));
    ProfScope ps(PROF_FRONTEND, s, (double)ntok * (2.0 * 16 * C + 16 * 8 + 8.0 * C), 4.0 * B * frames * 64 + 4.0 * ntok * C);
    static const bool use_mma = [] { const char* e = getenv("ARD_PATCH_EMBED_MMA"); return e == nullptr || atoi(e) != 0; }();
    if (use_mma && (C == 96 || C == 128)) {
        const unsigned gm = (unsigned)((ntok + 8 * 32 * PE_RUNS_PER_WARP - 1) / (8 * 32 * PE_RUNS_PER_WARP));
        if (C == 96)
            ARD_CUDA(enqueue_pdl(patch_embed_ln_mma_kernel<12>, dim3(gm), dim3(256), 0, s, logmel, clip_stride, frames, bn_scale, bn_shift, w, bias, gamma, beta, out, ntok));
        else
            ARD_CUDA(enqueue_pdl(patch_embed_ln_mma_kernel<16>, dim3(gm), dim3(256), 0, s, logmel, clip_stride, frames, bn_scale, bn_shift, w, bias, gamma, beta, out, ntok));
        return check_cuda(cudaGetLastError(), "patch_embed_ln launch");
    }
    if (C == 96)
        ARD_CUDA(enqueue_pdl(patch_embed_ln_kernel<3>, dim3(grid), dim3(256), 0, s, logmel, clip_stride, frames, bn_scale, bn_shift, w, bias, gamma, beta, out, ntok));
    else if (C == 128)
        ARD_CUDA(enqueue_pdl(patch_embed_ln_kernel<4>, dim3(grid), dim3(256), 0, s, logmel, clip_stride, frames, bn_scale, bn_shift, w, bias, gamma, beta, out, ntok));
    else
        return set_error(ARD_ERR_SHAPE, "patch_embed: unsupported embed_dim %d", C);
    return check_cuda(cudaGetLastError(), "patch_embed_ln launch");
}

}  // namespace ard
