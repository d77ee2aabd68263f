"""Stand-in for torchlibrosa 0.1.0 (unpinned third-party dependency of the reference,
CLAP/requirements.txt:3; source is NOT vendored under /root/reference).

Test infrastructure only: lets `oracle/refimport.py` import the reference's htsat.py
(CLAP/src/laion_clap/clap_module/htsat.py:21-22) in this container so golden vectors can be generated.
It restates the published algorithm of `torchlibrosa.stft.Spectrogram` / `LogmelFilterBank`
with the same module / parameter names so a reference state_dict round-trips:

  spectrogram_extractor.stft.conv_real.weight  [n_fft/2+1, 1, n_fft]   (cos * window)
  spectrogram_extractor.stft.conv_imag.weight  [n_fft/2+1, 1, n_fft]   (-sin * window)
  logmel_extractor.melW                        [n_fft/2+1, n_mels]
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def hann_periodic(n):
    # scipy.signal.get_window('hann', n, fftbins=True)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def hz_to_mel_slaney(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log_t = f >= min_log_hz
        mels = np.where(log_t, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)
    return mels


def mel_to_hz_slaney(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    log_t = m >= min_log_mel
    return np.where(log_t, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank_slaney(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel(htk=False, norm='slaney') restated: [n_mels, n_fft/2+1]."""
    n_freq = n_fft // 2 + 1
    fftfreqs = np.linspace(0, sr / 2.0, n_freq)
    mel_f = mel_to_hz_slaney(np.linspace(hz_to_mel_slaney(fmin), hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    W = np.zeros((n_mels, n_freq))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        W[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    W *= enorm[:, None]
    return W


class STFT(nn.Module):
    def __init__(self, n_fft, hop_length, win_length, window, center, pad_mode, freeze_parameters=True):
        super().__init__()
        assert window == "hann" and win_length == n_fft
        self.n_fft, self.hop_length, self.center, self.pad_mode = n_fft, hop_length, center, pad_mode
        out_channels = n_fft // 2 + 1
        self.conv_real = nn.Conv1d(1, out_channels, n_fft, stride=hop_length, bias=False)
        self.conv_imag = nn.Conv1d(1, out_channels, n_fft, stride=hop_length, bias=False)
        n = np.arange(n_fft)
        k = np.arange(out_channels)
        ang = 2.0 * np.pi * np.outer(k, n) / n_fft
        win = hann_periodic(n_fft)
        self.conv_real.weight.data = torch.tensor(np.cos(ang) * win[None, :], dtype=torch.float32)[:, None, :]
        self.conv_imag.weight.data = torch.tensor(-np.sin(ang) * win[None, :], dtype=torch.float32)[:, None, :]
        if freeze_parameters:
            for p in self.parameters():
                p.requires_grad = False

    def forward(self, x):
        x = x[:, None, :]
        if self.center:
            x = F.pad(x, pad=(self.n_fft // 2, self.n_fft // 2), mode=self.pad_mode)
        real = self.conv_real(x)
        imag = self.conv_imag(x)
        real = real[:, None, :, :].transpose(2, 3)
        imag = imag[:, None, :, :].transpose(2, 3)
        return real, imag


class Spectrogram(nn.Module):
    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
                 pad_mode="reflect", power=2.0, freeze_parameters=True):
        super().__init__()
        self.power = power
        self.stft = STFT(n_fft, hop_length, win_length, window, center, pad_mode, freeze_parameters)

    def forward(self, x):
        real, imag = self.stft(x)
        spec = real ** 2 + imag ** 2
        if self.power == 2.0:
            return spec
        return spec ** (self.power / 2.0)


class LogmelFilterBank(nn.Module):
    def __init__(self, sr=22050, n_fft=2048, n_mels=64, fmin=0.0, fmax=None, is_log=True, ref=1.0,
                 amin=1e-10, top_db=80.0, freeze_parameters=True):
        super().__init__()
        self.is_log, self.ref, self.amin, self.top_db = is_log, ref, amin, top_db
        if fmax is None:
            fmax = sr // 2
        melW = mel_filterbank_slaney(sr, n_fft, n_mels, fmin, fmax).T
        self.melW = nn.Parameter(torch.tensor(melW, dtype=torch.float32))
        if freeze_parameters:
            for p in self.parameters():
                p.requires_grad = False

    def forward(self, x):
        mel = torch.matmul(x, self.melW)
        if not self.is_log:
            return mel
        log_spec = 10.0 * torch.log10(torch.clamp(mel, min=self.amin, max=math.inf))
        log_spec = log_spec - 10.0 * math.log10(max(self.amin, self.ref))
        if self.top_db is not None:
            log_spec = torch.clamp(log_spec, min=log_spec.max().item() - self.top_db, max=math.inf)
        return log_spec
