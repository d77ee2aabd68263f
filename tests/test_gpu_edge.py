"""-m gpu: edge cases of the path the reference's own entry points accept: ragged clip lengths through the
featuriser, a batch of one, the numpy / int16 route of get_audio_embedding_from_data (hook.py:158-192), and the error
conventions of the C ABI mapped back to the exception types the reference raises."""
import ctypes as C

import numpy as np
import pytest
import torch

import gpu_checks as G
from audio_residual_b200 import lib as L
from oracle import htsat_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_ragged_clip_lengths_repeatpad_vs_oracle():
    """Clips of 0.4 s .. 10 s in one call: get_audio_features' repeatpad filling (data.py:466-486) then the encoder."""
    clap, sd, ores = G.make_encoder("tiny", residual=True)
    g = torch.Generator().manual_seed(11)
    lengths = [19200, 100001, 177777, 480000, 479999]
    clips = [(0.1 * torch.randn(n, generator=g)).clamp_(-1, 1) for n in lengths]
    with torch.no_grad():
        got = clap.get_audio_embedding_from_data([c.clone() for c in clips], use_tensor=True).float().cpu()
        wave = torch.stack([O.pad_clip(c, 480000, "repeatpad") for c in clips])
        ref = O.get_audio_embedding(wave, sd, O.CONFIGS["tiny"], ores)
    assert got.shape == (len(lengths), 512)
    assert _rel(got, ref) < G.TOL_BF16, _rel(got, ref)
    for mode in ("pad", "repeat"):
        with torch.no_grad():
            got = clap.get_audio_embedding_from_data([c.clone() for c in clips[:2]], use_tensor=True, data_fil=mode).float().cpu()
            ref = O.get_audio_embedding(torch.stack([O.pad_clip(c, 480000, mode) for c in clips[:2]]), sd, O.CONFIGS["tiny"], ores)
        assert _rel(got, ref) < G.TOL_BF16, (mode, _rel(got, ref))


def test_batch_of_one_and_numpy_int16_route():
    """use_tensor=False: numpy in, int16 round trip of the waveform (hook.py:177-179), numpy out; B = 1."""
    clap, sd, ores = G.make_encoder("tiny", residual=True)
    wave = G.W.make_clips(1, seed=21)
    wave[0, :7] = torch.tensor([1.0, -1.0, 1.3, -1.7, 0.0, 1.0 / 32767.0, -0.4 / 32767.0])     # values the quantiser clamps / truncates
    got = clap.get_audio_embedding_from_data(wave.numpy(), use_tensor=False)
    assert isinstance(got, np.ndarray) and got.shape == (1, 512) and got.dtype == np.float32
    with torch.no_grad():
        ref = O.get_audio_embedding(torch.from_numpy(O.int16_roundtrip_np(wave.numpy())), sd, O.CONFIGS["tiny"], ores)
    assert _rel(torch.from_numpy(got), ref) < G.TOL_BF16
    # the same clip inside a larger batch gives the same embedding bit for bit
    batch = torch.cat([wave, G.W.make_clips(6, seed=22)])
    got7 = clap.get_audio_embedding_from_data(batch.numpy(), use_tensor=False)
    assert np.array_equal(got7[0], got[0])


def test_abi_error_conventions():
    clap, sd, _ = G.make_encoder("tiny")
    enc = clap.model.audio_branch
    lib = L.load()
    with pytest.raises(AssertionError):                                  # wrong clip length (htsat.py:115 / :852 analogue)
        enc.encode(waveform=torch.zeros(2, 1000, device="cuda"))
    with pytest.raises(ValueError):                                      # src/residual.py:194-195
        L.check(lib.ard_set_block_residual(enc._handle(), 7, 0, None, None, 4, 96))
    a = L.ArdForwardArgs()
    a.B = 0
    assert lib.ard_encoder_forward(enc._handle(), C.byref(a), None) == L.ARD_ERR_SHAPE     # empty batch: an error code, no crash
    assert b"batch" in lib.ard_last_error()
    a.B = 2
    assert lib.ard_encoder_forward(enc._handle(), C.byref(a), None) == L.ARD_ERR_SHAPE     # no output buffer
    # unknown model size (htsat.py:1044-1045 raises RuntimeError; the shim passes exc=RuntimeError for this call)
    cfg = L.ArdConfig(embed_dim=100, depths=(C.c_int * 4)(2, 2, 6, 2), num_heads=(C.c_int * 4)(4, 8, 16, 32), joint_dim=512, enable_fusion=0)
    h = C.c_void_p()
    rc = lib.ard_create(C.byref(cfg), C.byref(h))
    assert rc == L.ARD_ERR_SHAPE and b"not found" in lib.ard_last_error()
    with pytest.raises(RuntimeError):
        L.check(rc, RuntimeError)
