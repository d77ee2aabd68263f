"""Synthetic, reproducible weights in the reference's state_dict layout.

There is no network in the build/bench environment, so the pretrained CLAP checkpoint (hook.py:91-112) cannot be
fetched; benchmarks and parity tests use random-init weights of the same architecture. The generator is keyed by
(seed, tensor name) through numpy's Generator so the same dict is reproduced bit-for-bit on any machine and can be
loaded both into this package and into the reference modules (CLAP/src/laion_clap/clap_module/htsat.py).

Scales are chosen so every path is exercised (unit-variance Linear outputs => non-uniform attention, non-trivial
LayerNorm/BatchNorm affines, non-zero biases), unlike the reference's trunc_normal(0.02)/zeros init
(htsat.py:761-768) under which softmax is ~uniform and biases vanish.
"""
import zlib

import numpy as np
import torch

CONFIGS = {
    # htsat.py:996-1027 + model_configs/HTSAT-{tiny,base}.json
    "tiny": dict(embed_dim=96, depths=(2, 2, 6, 2), num_heads=(4, 8, 16, 32), joint_dim=512),
    "base": dict(embed_dim=128, depths=(2, 2, 12, 2), num_heads=(4, 8, 16, 32), joint_dim=512),
}
N_FFT, HOP, MEL_BINS, SR, FMIN, FMAX, CLASS_NUM = 1024, 480, 64, 48000, 50, 14000, 527


def _rng(seed, key):
    return np.random.default_rng([seed, zlib.crc32(key.encode())])


def _randn(seed, key, *shape):
    return _rng(seed, key).standard_normal(shape, dtype=np.float32)


def hann_periodic(n):
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def _hz_to_mel(f, htk):
    f = np.asarray(f, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + f / 700.0)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_hz / f_sp + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m, htk):
    m = np.asarray(m, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    logstep = np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(htk=False, slaney_norm=True):
    """[513, 64] triangular mel filters. (htk=False, slaney_norm=True) is what torchlibrosa's LogmelFilterBank holds
    (librosa.filters.mel defaults); (htk=True, slaney_norm=False) is torchaudio's MelSpectrogram default used by the
    fusion featuriser (data.py:365-378)."""
    n_freq = N_FFT // 2 + 1
    fftfreqs = np.linspace(0, SR / 2.0, n_freq)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(FMIN, htk), _hz_to_mel(FMAX, htk), MEL_BINS + 2), htk)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    W = np.maximum(0, np.minimum(lower, upper))
    if slaney_norm:
        W = W * (2.0 / (mel_f[2:MEL_BINS + 2] - mel_f[:MEL_BINS]))[:, None]
    return np.ascontiguousarray(W.T)


def make_state_dict(model="tiny", seed=0, with_frontend=True):
    """dict[name -> float32 torch tensor] with the audio_branch keys un-prefixed (as `audio_branch.state_dict()`)
    plus the CLAP-level `audio_projection.{0,2}.{weight,bias}` keys (model.py:539-543)."""
    cfg = CONFIGS[model]
    C0, depths, heads = cfg["embed_dim"], cfg["depths"], cfg["num_heads"]
    sd = {}

    def lin(key, out_f, in_f, bias=True, gain=1.0):
        sd[key + ".weight"] = _randn(seed, key + ".weight", out_f, in_f) * np.float32(gain / np.sqrt(in_f))
        if bias:
            sd[key + ".bias"] = 0.1 * _randn(seed, key + ".bias", out_f)

    def norm(key, n):
        sd[key + ".weight"] = 1.0 + 0.1 * _randn(seed, key + ".weight", n)
        sd[key + ".bias"] = 0.1 * _randn(seed, key + ".bias", n)

    if with_frontend:
        n = np.arange(N_FFT)
        k = np.arange(N_FFT // 2 + 1)
        ang = 2.0 * np.pi * np.outer(k, n) / N_FFT
        win = hann_periodic(N_FFT)
        sd["spectrogram_extractor.stft.conv_real.weight"] = (np.cos(ang) * win[None]).astype(np.float32)[:, None, :]
        sd["spectrogram_extractor.stft.conv_imag.weight"] = (-np.sin(ang) * win[None]).astype(np.float32)[:, None, :]
        sd["logmel_extractor.melW"] = mel_filterbank().astype(np.float32)
    # log-mel of the synthetic clips (make_clips) sits near -11.5 dB with ~3.4 dB spread; these running stats keep
    # the normalised image O(1) while exercising all four BN terms.
    sd["bn0.weight"] = 1.0 + 0.1 * _randn(seed, "bn0.weight", MEL_BINS)
    sd["bn0.bias"] = 0.1 * _randn(seed, "bn0.bias", MEL_BINS)
    sd["bn0.running_mean"] = -11.5 + 0.5 * _randn(seed, "bn0.running_mean", MEL_BINS)
    sd["bn0.running_var"] = (11.5 * (1.0 + 0.2 * np.abs(_randn(seed, "bn0.running_var", MEL_BINS)))).astype(np.float32)

    sd["patch_embed.proj.weight"] = _randn(seed, "patch_embed.proj.weight", C0, 1, 4, 4) * np.float32(0.25)
    sd["patch_embed.proj.bias"] = 0.1 * _randn(seed, "patch_embed.proj.bias", C0)
    norm("patch_embed.norm", C0)
    for l, (d, nh) in enumerate(zip(depths, heads)):
        C = C0 << l
        for b in range(d):
            p = f"layers.{l}.blocks.{b}."
            norm(p + "norm1", C)
            sd[p + "attn.relative_position_bias_table"] = 0.5 * _randn(seed, p + "rpb", 225, nh)
            lin(p + "attn.qkv", 3 * C, C)
            lin(p + "attn.proj", C, C)
            norm(p + "norm2", C)
            lin(p + "mlp.fc1", 4 * C, C)
            lin(p + "mlp.fc2", C, 4 * C)
        if l < len(depths) - 1:
            norm(f"layers.{l}.downsample.norm", 4 * C)
            lin(f"layers.{l}.downsample.reduction", 2 * C, 4 * C, bias=False)
    NF = C0 << (len(depths) - 1)
    norm("norm", NF)
    sd["tscam_conv.weight"] = _randn(seed, "tscam_conv.weight", CLASS_NUM, NF, 2, 3) * np.float32(1.0 / np.sqrt(NF * 6))
    sd["tscam_conv.bias"] = 0.1 * _randn(seed, "tscam_conv.bias", CLASS_NUM)
    lin("audio_projection.0", cfg["joint_dim"], NF, gain=1.4)
    lin("audio_projection.2", cfg["joint_dim"], cfg["joint_dim"])
    return {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)) for k, v in sd.items()}


def make_pca(model="tiny", seed=0, layers=(0, 1, 2, 3)):
    """Per-layer ResiDual inputs in the reference pickle schema's dtype (src/residual.py:143-150):
    {layer: {"components": Q[D,D] float64 orthonormal, "mean": [D] float64}} and lambdas {layer: float32[D]}."""
    cfg = CONFIGS[model]
    pca, lam = {}, {}
    for l in layers:
        D = cfg["embed_dim"] << l
        q, r = np.linalg.qr(_rng(seed, f"pca{l}").standard_normal((D, D)))
        q = q * np.sign(np.diag(r))[None, :]
        pca[l] = {"components": np.ascontiguousarray(q.T), "mean": 0.1 * _rng(seed, f"pcamean{l}").standard_normal(D)}
        lam[l] = (1.0 + 0.1 * _randn(seed, f"lambda{l}", D)).astype(np.float32)
    return pca, lam


def make_clips(batch, seed=1234, n=480000):
    """Synthetic 10 s / 48 kHz clips: 0.1*randn plus three random sinusoids in 50..14000 Hz, clamped to [-1,1]."""
    out = np.empty((batch, n), dtype=np.float32)
    t = np.arange(n, dtype=np.float64) / SR
    for b in range(batch):
        r = _rng(seed, f"clip{b}")
        x = 0.1 * r.standard_normal(n, dtype=np.float32)
        for _ in range(3):
            f = r.uniform(50, 14000)
            x = x + (0.05 * r.uniform(0.2, 1.0) * np.sin(2 * np.pi * f * t + r.uniform(0, 6.28))).astype(np.float32)
        out[b] = np.clip(x, -1.0, 1.0)
    return torch.from_numpy(out)


def make_text_embeds(n_classes=50, dim=512, seed=7):
    e = _randn(seed, "text", n_classes, dim)
    e = e / np.linalg.norm(e, axis=1, keepdims=True)
    return torch.from_numpy(e.astype(np.float32))
