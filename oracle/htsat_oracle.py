"""CPU oracle: a functional restatement of the reference's HTSAT + ResiDual hot path.

TEST INFRASTRUCTURE — not product code. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module, and only as the checker / the reported CPU baseline.
The product path (audio_residual_b200/) never imports it and fails loudly if its CUDA library is missing.

The reference is pure PyTorch-Python (no native code), so the oracle is plain torch-on-CPU tensor code
(works in float32 and float64), written from the reference's arithmetic, each function citing the
reference file:line it follows (paths relative to /root/reference).

Parity status: PINNED. The reference's own tests hold no known-answer vectors for this path, so
the oracle is pinned against outputs of the reference code itself, imported in the build container by
oracle/refimport.py: oracle/make_golden.py runs both on identical seeded inputs/weights, asserts agreement, and
commits the reference's outputs as tests/golden/*.npz, which the `not gpu` tests re-check on every run.

Weights are passed as a flat dict using the reference's state_dict key names
(audio_branch keys un-prefixed, e.g. "layers.0.blocks.1.attn.qkv.weight").
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------- configs
# CLAP/src/laion_clap/clap_module/htsat.py:996-1027 (create_htsat_model) + model_configs/HTSAT-{tiny,base}.json
CONFIGS = {
    "tiny": dict(embed_dim=96, depths=(2, 2, 6, 2), num_heads=(4, 8, 16, 32), joint_dim=512),
    "base": dict(embed_dim=128, depths=(2, 2, 12, 2), num_heads=(4, 8, 16, 32), joint_dim=512),
}
WINDOW = 8          # htsat.py:1006 window_size=8
SPEC_SIZE = 256     # htsat.py:1000
MEL_BINS = 64
N_FFT = 1024
HOP = 480
CLIP_SAMPLES = 480000
FREQ_RATIO = SPEC_SIZE // MEL_BINS  # htsat.py:670
PATCH = 4
CLASS_NUM = 527
LN_EPS = 1e-5
BN_EPS = 1e-5


def num_features(cfg):
    return cfg["embed_dim"] * 2 ** (len(cfg["depths"]) - 1)


# ----------------------------------------------------------------------------------------------- featuriser
def quantize_tensor(x):
    """src/residual.py:210-212 (== src/analyze_attention.py:160-162): clamp, *32767 -> int16 (trunc) -> /32767."""
    x = torch.clamp(x, -1.0, 1.0)
    return (x * 32767.0).to(torch.int16).to(torch.float32) / 32767.0


def int16_roundtrip_np(x):
    """CLAP/src/laion_clap/training/data.py:93-99 (numpy twin used by hook.py:177 when use_tensor=False)."""
    x = np.clip(x, a_min=-1.0, a_max=1.0)
    return ((x * 32767.0).astype("int16") / 32767.0).astype("float32")


def pad_clip(wave, max_len=CLIP_SAMPLES, data_filling="repeatpad"):
    """data.py:469-496, the reachable (len <= max_len) branch of get_audio_features. wave: 1-D tensor."""
    n = wave.shape[0]
    if n > max_len:
        # data.py:467 calls np.random.integers which does not exist -> the reference crashes
        raise AttributeError("clips longer than max_len are unreachable in the reference (data.py:467)")
    if n == max_len:
        return wave
    if data_filling == "repeatpad":
        n_repeat = int(max_len / n)
        wave = wave.repeat(n_repeat)
        return F.pad(wave, (0, max_len - wave.shape[0]), mode="constant", value=0)
    if data_filling == "pad":
        return F.pad(wave, (0, max_len - n), mode="constant", value=0)
    if data_filling == "repeat":
        n_repeat = int(max_len / n)
        return wave.repeat(n_repeat + 1)[:max_len]
    raise NotImplementedError(f"data_filling {data_filling} not implemented")


# ----------------------------------------------------------------------------------------------- front end
def stft_power(wave, sd):
    """torchlibrosa Spectrogram as constructed at htsat.py:681-683 and called at :898.
    wave [B,T] -> [B,1,frames,513]: reflect-pad n_fft/2, conv1d with the (window*DFT) kernels, re^2+im^2."""
    x = wave[:, None, :]
    x = F.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    wr = sd["spectrogram_extractor.stft.conv_real.weight"].to(x.dtype)
    wi = sd["spectrogram_extractor.stft.conv_imag.weight"].to(x.dtype)
    real = F.conv1d(x, wr, stride=HOP)
    imag = F.conv1d(x, wi, stride=HOP)
    spec = real ** 2 + imag ** 2                      # [B,513,frames]
    return spec[:, None].transpose(2, 3)


def logmel(spec, sd):
    """torchlibrosa LogmelFilterBank(ref=1.0, amin=1e-10, top_db=None), htsat.py:685-687, :899."""
    mel = torch.matmul(spec, sd["logmel_extractor.melW"].to(spec.dtype))
    out = 10.0 * torch.log10(torch.clamp(mel, min=1e-10))
    return out - 10.0 * math.log10(max(1e-10, 1.0))


def bn0_eval(x, sd):
    """htsat.py:900-902: transpose(1,3) -> BatchNorm2d(64) in eval mode -> transpose back. x [B,1,T,64]."""
    dt = x.dtype
    rm, rv = sd["bn0.running_mean"].to(dt), sd["bn0.running_var"].to(dt)
    g, b = sd["bn0.weight"].to(dt), sd["bn0.bias"].to(dt)
    return (x - rm) / torch.sqrt(rv + BN_EPS) * g + b


def reshape_wav2img(x):
    """htsat.py:848-863. x [B,1,T,F] -> [B,1,256,256]; bicubic (align_corners=True) along T to 1024, then the
    1024 frames are folded into 4 frequency-stacked quarters: img[b,0,r*64+f,t] = x[b,0,r*256+t,f]."""
    B, C, T, Fq = x.shape
    target_T = SPEC_SIZE * FREQ_RATIO
    target_F = SPEC_SIZE // FREQ_RATIO
    assert T <= target_T and Fq <= target_F, "the wav size should less than or equal to the swin input size"
    if T < target_T:
        x = F.interpolate(x, (target_T, x.shape[3]), mode="bicubic", align_corners=True)
    if Fq < target_F:
        x = F.interpolate(x, (x.shape[2], target_F), mode="bicubic", align_corners=True)
    x = x.permute(0, 1, 3, 2).contiguous()
    x = x.reshape(x.shape[0], x.shape[1], x.shape[2], FREQ_RATIO, x.shape[3] // FREQ_RATIO)
    x = x.permute(0, 1, 3, 2, 4).contiguous()
    return x.reshape(x.shape[0], x.shape[1], x.shape[2] * x.shape[3], x.shape[4])


def patch_embed(img, sd):
    """PatchEmbed.forward, htsat.py:108-144 (non-fusion branch and the fusion branch with longer_idx=[] are the
    same arithmetic: `proj` conv 4x4 stride 4 on channel 0, flatten, LayerNorm)."""
    dt = img.dtype
    B, C, H, W = img.shape
    assert H == SPEC_SIZE and W == SPEC_SIZE, \
        f"Input image size ({H}*{W}) doesn't match model ({SPEC_SIZE}*{SPEC_SIZE})."
    x = F.conv2d(img[:, 0:1], sd["patch_embed.proj.weight"].to(dt), sd["patch_embed.proj.bias"].to(dt), stride=PATCH)
    x = x.flatten(2).transpose(1, 2)
    Cc = x.shape[-1]
    return F.layer_norm(x, (Cc,), sd["patch_embed.norm.weight"].to(dt), sd["patch_embed.norm.bias"].to(dt), LN_EPS)


# ----------------------------------------------------------------------------------------------- swin pieces
def window_partition(x, ws=WINDOW):
    """htsat.py:248-259"""
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def window_reverse(windows, ws, H, W):
    """htsat.py:262-275"""
    B = int(windows.shape[0] / (H * W / ws / ws))
    x = windows.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def relative_position_index(ws=WINDOW):
    """htsat.py:301-316"""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def shift_attn_mask(H, W, ws=WINDOW, shift=WINDOW // 2):
    """htsat.py:414-437: 9 region labels -> per-window additive mask in {0, -100}."""
    img_mask = torch.zeros((1, H, W, 1))
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img_mask[:, h, w, :] = cnt
            cnt += 1
    mw = window_partition(img_mask, ws).view(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, float(-100.0)).masked_fill(am == 0, float(0.0))


def window_attention(x, sd, p, num_heads, mask, taps=None):
    """WindowAttention.forward, htsat.py:326-357. x [B_,64,C] -> (out [B_,64,C], attn [B_,nH,64,64]).
    `taps` (a list) receives the per-head `attn @ v` temporary [B_, nH, 64, hd] of htsat.py:354 (before its transpose)."""
    dt = x.dtype
    B_, N, C = x.shape
    hd = C // num_heads
    qkv = F.linear(x, sd[p + "attn.qkv.weight"].to(dt), sd[p + "attn.qkv.bias"].to(dt))
    qkv = qkv.reshape(B_, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (hd ** -0.5)
    attn = q @ k.transpose(-2, -1)
    idx = relative_position_index()
    bias = sd[p + "attn.relative_position_bias_table"].to(dt)[idx.view(-1)].view(N, N, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, num_heads, N, N) + mask.to(dt).unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, num_heads, N, N)
    attn = torch.softmax(attn, dim=-1)
    head_out = attn @ v
    if taps is not None:
        taps.append(head_out)
    out = head_out.transpose(1, 2).reshape(B_, N, C)
    out = F.linear(out, sd[p + "attn.proj.weight"].to(dt), sd[p + "attn.proj.bias"].to(dt))
    return out, attn


def mlp(x, sd, p):
    """Mlp.forward htsat.py:158-164 (exact-erf GELU, dropout p=0)."""
    dt = x.dtype
    x = F.linear(x, sd[p + "mlp.fc1.weight"].to(dt), sd[p + "mlp.fc1.bias"].to(dt))
    x = F.gelu(x)
    return F.linear(x, sd[p + "mlp.fc2.weight"].to(dt), sd[p + "mlp.fc2.bias"].to(dt))


def residual_apply(x, mean, basis, lam):
    """ResiDual.forward, src/residual.py:29-42: ((x - mean) @ basis.T * lam) @ basis; the mean is NOT re-added."""
    xc = x - mean
    return torch.matmul(torch.matmul(xc, basis.T) * lam, basis)


def swin_block(x, sd, p, H, W, num_heads, shift, residual=None, taps=None):
    """SwinTransformerBlock.forward htsat.py:439-482, or, when `residual=(mean,basis,lam)` is given, the patched
    forward of src/residual.py:58-98 including its doubled shortcut/FFN.
    Eval mode: DropPath is the identity. Returns (x, attn, residual_x)."""
    dt = x.dtype
    B, L, C = x.shape
    ws = WINDOW
    if min(H, W) <= ws:            # htsat.py:393-396
        shift = 0
        ws = min(H, W)
    shortcut = x
    xn = F.layer_norm(x, (C,), sd[p + "norm1.weight"].to(dt), sd[p + "norm1.bias"].to(dt), LN_EPS).view(B, H, W, C)
    if shift > 0:
        xn = torch.roll(xn, shifts=(-shift, -shift), dims=(1, 2))
        mask = shift_attn_mask(H, W, ws, shift)
    else:
        mask = None
    xw = window_partition(xn, ws).view(-1, ws * ws, C)
    aw, attn = window_attention(xw, sd, p, num_heads, mask, taps)
    xs = window_reverse(aw.view(-1, ws, ws, C), ws, H, W)
    if shift > 0:
        xs = torch.roll(xs, shifts=(shift, shift), dims=(1, 2))
    residual_x = xs.view(B, H * W, C)

    def n2(t):
        return F.layer_norm(t, (C,), sd[p + "norm2.weight"].to(dt), sd[p + "norm2.bias"].to(dt), LN_EPS)

    if residual is None:
        x = shortcut + residual_x                       # htsat.py:479
        x = x + mlp(n2(x), sd, p)                       # htsat.py:480
        return x, attn, residual_x
    mean, basis, lam = residual
    residual_x = residual_apply(residual_x, mean.to(dt), basis.to(dt), lam.to(dt))   # src/residual.py:88-89
    x = shortcut + residual_x                           # src/residual.py:92
    x = x + mlp(n2(x), sd, p)                           # src/residual.py:93
    x = shortcut + x                                    # src/residual.py:95
    x = x + mlp(n2(x), sd, p)                           # src/residual.py:96
    return x, attn, residual_x


def patch_merging(x, sd, p, H, W):
    """PatchMerging.forward htsat.py:505-526."""
    dt = x.dtype
    B, L, C = x.shape
    assert L == H * W, "input feature has wrong size"
    assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
    x = x.view(B, H, W, C)
    x = torch.cat([x[:, 0::2, 0::2, :], x[:, 1::2, 0::2, :], x[:, 0::2, 1::2, :], x[:, 1::2, 1::2, :]], -1)
    x = x.view(B, -1, 4 * C)
    x = F.layer_norm(x, (4 * C,), sd[p + "downsample.norm.weight"].to(dt), sd[p + "downsample.norm.bias"].to(dt), LN_EPS)
    return F.linear(x, sd[p + "downsample.reduction.weight"].to(dt))


def basic_layer(x, sd, l, cfg, H, W, residual=None, taps=None):
    """BasicLayer.forward htsat.py:580-597 in eval mode: attention maps are stacked and averaged over the layer's
    blocks, residual_x tensors are concatenated along tokens."""
    attns, ress = [], []
    for b in range(cfg["depths"][l]):
        shift = 0 if b % 2 == 0 else WINDOW // 2
        x, a, r = swin_block(x, sd, f"layers.{l}.blocks.{b}.", H, W, cfg["num_heads"][l], shift, residual, taps)
        attns.append(a.unsqueeze(0))
        ress.append(r)
    if l < len(cfg["depths"]) - 1:
        x = patch_merging(x, sd, f"layers.{l}.", H, W)
    attn = torch.mean(torch.cat(attns, dim=0), dim=0)
    return x, attn, torch.cat(ress, dim=1)


def forward_features(img, sd, cfg, residuals=None, head_outputs=False):
    """HTSAT_Swin_Transformer.forward_features htsat.py:779-834. residuals: {layer: (mean, basis, lam)}.
    head_outputs: add key "head_outputs", per layer [depth, B*nW, nH, 64, hd] (every block's per-head attn @ v, htsat.py:354)."""
    dt = img.dtype
    residuals = residuals or {}
    frames_num = img.shape[2]
    x = patch_embed(img, sd)
    H = W = SPEC_SIZE // PATCH
    attns, ress, heads = [], [], []
    for l in range(len(cfg["depths"])):
        taps = [] if head_outputs else None
        x, a, r = basic_layer(x, sd, l, cfg, H >> l, W >> l, residuals.get(l), taps)
        attns.append(a)
        ress.append(r)
        if head_outputs:
            heads.append(torch.stack(taps, dim=0))
    Cn = x.shape[-1]
    x = F.layer_norm(x, (Cn,), sd["norm.weight"].to(dt), sd["norm.bias"].to(dt), LN_EPS)
    B, N, C = x.shape
    nl = len(cfg["depths"])
    SF = frames_num // (2 ** (nl - 1)) // PATCH
    ST = frames_num // (2 ** (nl - 1)) // PATCH
    x = x.permute(0, 2, 1).contiguous().reshape(B, C, SF, ST)
    c_freq_bin = SF // FREQ_RATIO
    x = x.reshape(B, C, SF // c_freq_bin, c_freq_bin, ST)
    x = x.permute(0, 1, 3, 2, 4).contiguous().reshape(B, C, c_freq_bin, -1)
    fine = torch.mean(x, dim=2)
    fine = interpolate_repeat(fine.permute(0, 2, 1).contiguous(), 8 * PATCH)
    latent = torch.flatten(x, 2).mean(dim=-1)                                   # AdaptiveAvgPool1d(1)
    y = F.conv2d(x, sd["tscam_conv.weight"].to(dt), sd["tscam_conv.bias"].to(dt), padding=(0, 1))
    y = torch.flatten(y, 2)
    fpx = interpolate_repeat(torch.sigmoid(y).permute(0, 2, 1).contiguous(), 8 * PATCH)
    clip = torch.sigmoid(y.mean(dim=-1))
    out = {"framewise_output": fpx, "clipwise_output": clip, "fine_grained_embedding": fine,
           "embedding": latent, "layers_attention": attns, "layers_residuals": ress}
    if head_outputs:
        out["head_outputs"] = heads
    return out


def interpolate_repeat(x, ratio):
    """clap_module/utils.py:209-224"""
    B, T, C = x.shape
    return x[:, :, None, :].repeat(1, 1, ratio, 1).reshape(B, T * ratio, C)


def htsat_forward(inputs, sd, cfg, residuals=None, enable_fusion=False, head_outputs=False):
    """HTSAT_Swin_Transformer.forward htsat.py:881-994 restricted to the reachable eval-mode routes:
    non-fusion (waveform -> STFT -> logmel -> bn0 -> img) and fusion with no `longer` clip (mel_fusion -> bn0 -> img,
    longer_idx=[] so only channel 0 reaches patch_embed.proj, htsat.py:883-894 / :108-134)."""
    if enable_fusion:
        x = inputs["mel_fusion"]                       # [B,4,T,64]
        x = bn0_eval(x, sd)
        x = reshape_wav2img(x)
        return forward_features(x, sd, cfg, residuals, head_outputs)
    x = stft_power(inputs["waveform"], sd)
    x = logmel(x, sd)
    x = bn0_eval(x, sd)
    x = reshape_wav2img(x)
    return forward_features(x, sd, cfg, residuals, head_outputs)


def audio_projection(emb, sd):
    """CLAP.audio_projection (model.py:539-543: Linear, ReLU, Linear) + F.normalize (model.py:739-741).
    Keys as in the CLAP state_dict: audio_projection.{0,2}.{weight,bias}."""
    dt = emb.dtype
    h = F.relu(F.linear(emb, sd["audio_projection.0.weight"].to(dt), sd["audio_projection.0.bias"].to(dt)))
    h = F.linear(h, sd["audio_projection.2.weight"].to(dt), sd["audio_projection.2.bias"].to(dt))
    return F.normalize(h, dim=-1)


def get_audio_embedding(wave, sd, cfg, residuals=None):
    """CLAP.get_audio_embedding model.py:720-742 on a batched waveform tensor."""
    out = htsat_forward({"waveform": wave}, sd, cfg, residuals)
    return audio_projection(out["embedding"], sd)


def fusion_mel(wave, htk_fb, window):
    """get_mel, data.py:363-399: torchaudio MelSpectrogram(n_fft=1024, hop=480, center, reflect, power=2,
    norm=None, mel_scale='htk', n_mels=64, f 50..14000) + AmplitudeToDB(top_db=None) -> [frames, 64].
    htk_fb [513,64] is the htk filterbank (torchaudio.functional.melscale_fbanks); window = periodic hann."""
    x = F.pad(wave[None, None, :], (N_FFT // 2, N_FFT // 2), mode="reflect")[0, 0]
    frames = x.unfold(0, N_FFT, HOP) * window
    spec = torch.fft.rfft(frames, dim=-1)
    power = spec.real ** 2 + spec.imag ** 2
    mel = power @ htk_fb
    return 10.0 * torch.log10(torch.clamp(mel, min=1e-10))


# ----------------------------------------------------------------------------------------------- training step
def zero_shot_loss(wave, labels, text_embeds, sd, cfg, residuals):
    """train_one_epoch_zero_shot src/training.py:12-41: similarities = emb @ text.T (no logit scale),
    CrossEntropyLoss(mean)."""
    emb = get_audio_embedding(wave, sd, cfg, residuals)
    sims = emb @ text_embeds.T.to(emb.dtype)
    return F.cross_entropy(sims, labels), sims


def linear_probe_loss(wave, labels, W, b, sd, cfg, residuals=None):
    """HTSATLinearClassifier.forward src/linear.py:27-32 + CE (src/linear.py:43-44)."""
    emb = get_audio_embedding(wave, sd, cfg, residuals)
    logits = F.linear(emb, W.to(emb.dtype), b.to(emb.dtype))
    return F.cross_entropy(logits, labels), logits


# ----------------------------------------------------------------------------------------------- PCA statistics
def pca_from_moments(n, s1, s2):
    """What sklearn IncrementalPCA(n_components=None) converges to on full-rank data (src/residual.py:110,138,
    143-150), written from the sufficient statistics n, s1 = sum x [D], s2 = sum x x^T [D,D] (float64):
    mean, eigen-decomposition of the ddof=1 covariance, components sorted by decreasing variance, sign fixed by
    sklearn's svd_flip(u_based_decision=False): the largest-|entry| of each component row is positive."""
    s1 = np.asarray(s1, dtype=np.float64)
    s2 = np.asarray(s2, dtype=np.float64)
    mean = s1 / n
    cov = (s2 - n * np.outer(mean, mean)) / (n - 1)
    cov = 0.5 * (cov + cov.T)
    w, v = np.linalg.eigh(cov)
    order = np.argsort(w)[::-1]
    w = np.maximum(w[order], 0.0)
    comps = v[:, order].T
    idx = np.argmax(np.abs(comps), axis=1)
    signs = np.sign(comps[np.arange(comps.shape[0]), idx])
    signs[signs == 0] = 1.0
    comps = comps * signs[:, None]
    total = w.sum()
    return {"components": comps, "mean": mean, "explained_variance": w,
            "explained_variance_ratio": w / total, "n_components": comps.shape[0],
            "input_dim": comps.shape[1], "num_samples": int(n)}


def spectrum_summaries(explained_variance, explained_variance_ratio):
    """save_pca_results_on_file src/analyze_attention.py:80-83: intrinsic dim = #(cumsum(ratio) < 0.99) + 1,
    participation ratio = (sum v)^2 / sum v^2."""
    cumsum = np.cumsum(explained_variance_ratio)
    intrinsic_dim = int((cumsum < 0.99).sum() + 1)
    pr = float((explained_variance.sum() ** 2) / np.sum(explained_variance ** 2))
    return pr, intrinsic_dim
