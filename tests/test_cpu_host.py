"""`not gpu`: host-side logic of the drop-in (module tree / state_dict parity, ResiDual injection semantics, featuriser,
PCA artefact schema, CSV schema) and the C-ABI library's exported surface. No compute calls: there is no GPU here."""
import copy
import ctypes
import json
import os
import pickle

import numpy as np
import pytest
import torch

from audio_residual_b200 import lib as L
from audio_residual_b200 import weights as W
from audio_residual_b200.analyze_attention import load_pca_csv_results, save_pca_results_on_file
from audio_residual_b200.clap import CLAP_Module, batch_features, float32_to_int16, get_audio_features, int16_to_float32
from audio_residual_b200.residual import (ResiDual, load_residual, patch_block_with_residual, pca_from_moments, quantize_tensor,
                                          setup_residual_htsat)
from oracle import htsat_oracle as O

from gpu_checks import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------------------------------- C ABI surface
def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    if not os.path.exists(L.LIB_PATH):
        ge.build()
    lib = L.load(check_symbols=True)
    names = L.declared_symbols()
    assert len(names) >= 20 and "ard_encoder_forward" in names and "ard_gemm_bf16" in names
    for n in names:
        assert hasattr(lib, n), n
    assert lib.ard_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device ard_create must fail loudly (ARD_ERR_CUDA -> RuntimeError), never fall back."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load()
    cfg = L.ArdConfig(96, (ctypes.c_int * 4)(2, 2, 6, 2), (ctypes.c_int * 4)(4, 8, 16, 32), 512, 0)
    h = ctypes.c_void_p()
    rc = lib.ard_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == L.ARD_ERR_CUDA
    with pytest.raises(RuntimeError):
        L.check(rc)
    assert b"no CPU fallback" in lib.ard_last_error()
    m = CLAP_Module(device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.get_audio_embedding_from_data(torch.zeros(1, 480000), use_tensor=True)


def test_create_rejects_unknown_model():
    lib = L.load()
    cfg = L.ArdConfig(100, (ctypes.c_int * 4)(2, 2, 6, 2), (ctypes.c_int * 4)(4, 8, 16, 32), 512, 0)
    h = ctypes.c_void_p()
    assert lib.ard_create(ctypes.byref(cfg), ctypes.byref(h)) == L.ARD_ERR_SHAPE    # htsat.py:1044-1045 RuntimeError in the shim
    assert lib.ard_set_weight(None, b"x", None, 0) == L.ARD_ERR_SHAPE


# ---------------------------------------------------------------------------------------------------- module tree
@pytest.mark.parametrize("name,fusion", [("tiny", False), ("base", True)])
def test_state_dict_matches_reference_keys(name, fusion):
    ref = json.load(open(os.path.join(GOLDEN, "reference_state_dict.json")))[f"{name}{'_fusion' if fusion else ''}"]
    m = CLAP_Module(enable_fusion=fusion, device="cpu", amodel=f"HTSAT-{name}")
    own = {k: list(v.shape) for k, v in m.model.audio_branch.state_dict().items()}
    own.update({"audio_projection." + k: list(v.shape) for k, v in m.model.audio_projection.state_dict().items()})
    # the aff_2d fusion branch (mel_conv2d + AFF) is unreachable in the reference and not instantiated
    ref = {k: v for k, v in ref.items() if "mel_conv2d" not in k and "fusion_model" not in k}
    assert set(own) == set(ref), (set(own) ^ set(ref))
    assert all(own[k] == ref[k] for k in ref)
    sd = W.make_state_dict(name, seed=0)
    for k, v in sd.items():
        assert list(v.shape) == ref[k], k


def test_blocks_addressable_and_deepcopy_relinks():
    m = CLAP_Module(device="cpu")
    enc = m.model.audio_branch
    assert len(enc.layers) == 4 and [len(l.blocks) for l in enc.layers] == [2, 2, 6, 2]
    assert enc.layers[0].blocks[1].shift_size == 4 and enc.layers[3].blocks[1].shift_size == 0     # htsat.py:393-396
    assert enc.layers[0].blocks[1].attn_mask.shape == (64, 64, 64) and enc.layers[3].blocks[1].attn_mask is None
    c = copy.deepcopy(enc)
    assert c.layers[2].blocks[3]._encoder() is c and enc.layers[2].blocks[3]._encoder() is enc
    assert c._hb is not enc._hb


# ---------------------------------------------------------------------------------------------------- ResiDual API
def _pca_files(tmp_path, layers=(0, 1, 2, 3)):
    pca, _ = W.make_pca("tiny", seed=0, layers=layers)
    files = {}
    for l, d in pca.items():
        files[l] = str(tmp_path / f"layer_{l}")
        with open(files[l], "wb") as f:
            pickle.dump({"components": d["components"], "mean": d["mean"]}, f)
    return files, pca


def test_setup_residual_htsat_semantics(tmp_path):
    files, pca = _pca_files(tmp_path)
    m = CLAP_Module(device="cpu")
    enc = m.model.audio_branch
    new, residuals = setup_residual_htsat(enc, files, [0, 2])
    assert new is not enc and set(residuals) == {0, 2}
    assert all(b._residual is None for l in enc.layers for b in l.blocks)                # original untouched (deepcopy)
    assert all(b._residual is residuals[0] for b in new.layers[0].blocks)                # one ResiDual per layer, shared (Q3)
    assert all(b._residual is residuals[2] for b in new.layers[2].blocks)
    assert all(b._residual is None for b in new.layers[1].blocks)
    assert not any(p.requires_grad for p in new.parameters())                            # encoder frozen
    assert not any("learnable" in n for n, _ in new.named_parameters())                  # not registered (Q4)
    r = residuals[0]
    assert r.learnable.requires_grad and torch.equal(r.learnable.data, torch.ones(96))
    assert r.basis.shape == (96, 96) and r.mean.shape == (96,) and r.basis.dtype == torch.float32
    assert np.allclose(r.basis.numpy(), pca[0]["components"].astype(np.float32))
    with pytest.raises(ValueError, match="out of range"):
        setup_residual_htsat(enc, files, [4])
    m.model.audio_branch = new                                                           # src/training.py:103
    assert new._projection is m.model.audio_projection


def test_residual_module_surface():
    basis = torch.linalg.qr(torch.randn(32, 32))[0].T.contiguous()
    r = ResiDual(basis, torch.zeros(32), n_components=8)
    assert r.n_components == 8 and r.basis.shape == (8, 32) and r.learnable.shape == (8,)     # rows are sliced (Q11)
    assert set(dict(r.named_buffers())) == {"mean", "basis"}
    with pytest.raises(RuntimeError, match="CUDA"):
        r(torch.zeros(1, 2, 32))
    blk = CLAP_Module(device="cpu").model.audio_branch.layers[0].blocks[0]
    patch_block_with_residual(blk, r)
    assert blk._residual is r and "_residual" not in dict(blk.named_modules())


def test_load_residual_reads_reference_fixture_schema(tmp_path):
    files, pca = _pca_files(tmp_path, layers=(1,))
    r = load_residual(files[1])
    assert r.basis.shape == (192, 192) and torch.allclose(r.basis @ r.basis.T, torch.eye(192), atol=1e-5)


# ---------------------------------------------------------------------------------------------------- featuriser
def test_featuriser_matches_oracle():
    rng = np.random.default_rng(0)
    clips = [torch.from_numpy(rng.standard_normal(n).astype(np.float32)) for n in (1000, 4800, 12000)]
    for mode in ("repeatpad", "pad", "repeat"):
        got = batch_features(clips, 12000, mode)
        for i, c in enumerate(clips):
            assert torch.equal(got[i], O.pad_clip(c, 12000, mode))
        s = get_audio_features({}, clips[0], 12000, "rand_trunc", mode, {})
        assert torch.equal(s["waveform"], got[0]) and s["longer"].tolist() == [False]
    with pytest.raises(NotImplementedError):
        batch_features(clips, 12000, "bogus")
    with pytest.raises(NotImplementedError):
        get_audio_features({}, clips[0], 12000, "bogus", "pad", {})
    with pytest.raises(AttributeError):                                    # reference crashes for > max_len (Q9)
        get_audio_features({}, torch.zeros(13000), 12000, "rand_trunc", "pad", {})
    x = rng.uniform(-1.2, 1.2, 100).astype(np.float32)
    assert np.array_equal(int16_to_float32(float32_to_int16(x)), O.int16_roundtrip_np(x))
    assert torch.equal(quantize_tensor(torch.from_numpy(x)), O.quantize_tensor(torch.from_numpy(x)))


# ---------------------------------------------------------------------------------------------------- PCA artefacts
def test_pca_from_moments_matches_oracle_and_schema():
    g = np.load(os.path.join(GOLDEN, "pca_moments.npz"))
    got = pca_from_moments(int(g["n"]), g["s1"], g["s2"])
    ref = O.pca_from_moments(int(g["n"]), g["s1"], g["s2"])
    assert set(got) == {"components", "mean", "explained_variance", "explained_variance_ratio", "n_components", "input_dim",
                        "num_samples"}                                    # src/residual.py:143-150
    for k in ("components", "mean", "explained_variance", "explained_variance_ratio"):
        assert np.allclose(got[k], ref[k])
    assert np.allclose(got["components"], g["components"], atol=5e-6)       # == sklearn IncrementalPCA incl. sign rule
    trunc = pca_from_moments(int(g["n"]), g["s1"], g["s2"], n_components=10)
    assert trunc["components"].shape == (10, 96) and trunc["n_components"] == 10


def test_csv_schema_roundtrip(tmp_path):
    class P:
        pass
    models = {0: {}, 1: {}}
    rng = np.random.default_rng(1)
    for l in models:
        for h in range(2):
            p = P()
            v = np.sort(rng.uniform(0.1, 2, 16))[::-1]
            p.explained_variance_, p.explained_variance_ratio_ = v, v / v.sum()
            models[l][h] = p
    path = save_pca_results_on_file(str(tmp_path), "ESC50", 0, models)
    header = open(path).readline().strip().split(",")
    assert header == ["layer", "head", "component_index", "explained_variance", "explained_variance_ratio", "participation_ratio",
                      "intrinsic_dim"]                                     # src/analyze_attention.py:70-76
    back = load_pca_csv_results(path)
    v = models[1][1].explained_variance_
    assert np.allclose(back[(1, 1)]["explained_variance"], v)
    pr, idim = O.spectrum_summaries(v, v / v.sum())
    assert abs(back[(1, 1)]["participation_ratio"] - pr) < 1e-9 and back[(1, 1)]["intrinsic_dim"] == idim


def test_real_reference_pca_fixture_schema():
    """The reference ships real PCA pickles (residual_pca/ESC50/layer_*_evalfold_*); when mounted, they must load."""
    p = "/root/reference/residual_pca/ESC50/layer_0_evalfold_0"
    if not os.path.exists(p):
        pytest.skip("reference fixtures not mounted")
    r = load_residual(p)
    assert r.basis.shape == (96, 96) and r.mean.shape == (96,)
    assert torch.allclose(r.basis @ r.basis.T, torch.eye(96), atol=1e-4)


# ---------------------------------------------------------------------------------------------------- checkpoints
def test_load_ckpt_reference_layout(tmp_path):
    """hook.py:75-119: a LAION-CLAP style checkpoint ({"state_dict": {"module.audio_branch....": ...}}) loads into the mirror;
    text-tower tensors are ignored, a missing audio tensor raises like nn.Module.load_state_dict does."""
    sd = W.make_state_dict("tiny", seed=3)
    ck = {}
    for k, v in sd.items():
        ck["module." + (k if k.startswith("audio_projection.") else "audio_branch." + k)] = v
    for k, v in CLAP_Module(device="cpu").model.audio_branch.state_dict().items():   # buffers + the unused `head` Linear a real checkpoint carries
        ck.setdefault("module.audio_branch." + k, v)
    ck["module.text_branch.embeddings.word_embeddings.weight"] = torch.zeros(4, 4)
    ck["module.logit_scale_a"] = torch.tensor(1.0)
    path = tmp_path / "ckpt.pt"
    torch.save({"state_dict": ck, "epoch": 1}, path)
    m = CLAP_Module(device="cpu").load_ckpt(str(path), verbose=False)
    got = m.model.audio_branch.state_dict()
    for k in ("layers.2.blocks.3.attn.qkv.weight", "bn0.running_var", "patch_embed.proj.weight", "logmel_extractor.melW"):
        assert torch.equal(got[k], torch.as_tensor(sd[k])), k
    assert torch.equal(m.model.audio_projection.state_dict()["2.weight"], torch.as_tensor(sd["audio_projection.2.weight"]))
    del ck["module.audio_branch.layers.0.blocks.0.norm1.weight"]
    with pytest.raises(RuntimeError, match="Missing key"):
        CLAP_Module(device="cpu").load_ckpt({"state_dict": ck}, verbose=False)
    with pytest.raises(RuntimeError, match="download"):
        CLAP_Module(device="cpu").load_ckpt(None)


def test_lambda_save_restore(tmp_path):
    from audio_residual_b200.residual import inject_residuals, load_lambdas, save_lambdas
    m = CLAP_Module(device="cpu")
    pca, lam = W.make_pca("tiny", seed=0)
    res = inject_residuals(m.model.audio_branch, pca, lam)
    save_lambdas(res, tmp_path / "lam.pt")
    for r in res.values():
        with torch.no_grad():
            r.learnable.fill_(1.0)
    load_lambdas(res, tmp_path / "lam.pt")
    for l, r in res.items():
        assert torch.allclose(r.learnable.detach(), torch.as_tensor(lam[l]))


def test_host_pipeline_chunk_schedule():
    """_chunk_bounds (clap.py): contiguous cover of the batch, a small first chunk (the only exposed copy), every later copy
    sized to fit under the previous chunk's encode (next <= 22.8 + 1.156 x current, measured model), no tiny tail."""
    m = CLAP_Module.__new__(CLAP_Module)          # the schedule needs no model state
    for N in (65, 100, 137, 256, 512, 1000, 4096):
        b = CLAP_Module._chunk_bounds(m, N)
        assert b[0][0] == 0 and b[-1][1] == N and all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
        sizes = [hi - lo for lo, hi in b]
        assert all(s > 0 for s in sizes) and sizes[0] <= 24 and max(sizes) <= 256 + 256 // 3
        for cur, nxt in zip(sizes[:-2], sizes[1:-1]):            # the last chunk may absorb a short tail
            assert nxt <= 22.8 + 1.156 * cur + 1, (N, sizes)
        assert sizes[-1] >= min(8, N) or len(sizes) == 1, (N, sizes)


def test_int16_dequantisation_is_exact_for_every_sample_value():
    """The device featuriser divides int16 samples by 32767 in fp32; numpy's int16_to_float32 (data.py:93-94) divides in float64
    and rounds once. The two agree for all 65536 inputs, so the int16 PCM transport is bit-identical to the float route."""
    q = np.arange(-32768, 32768, dtype=np.int64).astype(np.int16)
    assert np.array_equal((q / 32767.0).astype(np.float32), q.astype(np.float32) / np.float32(32767.0))


def test_batch_features_host_path_int16_and_quantize():
    from audio_residual_b200.clap import batch_features
    from oracle import htsat_oracle as O
    g = torch.Generator().manual_seed(3)
    clips = [(0.5 * torch.randn(n, generator=g)).clamp_(-1.3, 1.3) for n in (5, 33, 100)]
    ref = torch.stack([O.pad_clip(c, 100, "repeatpad") for c in clips])
    assert torch.equal(batch_features(clips, 100, "repeatpad"), ref)
    assert torch.equal(batch_features(clips, 100, "repeatpad", quantize=True), O.quantize_tensor(ref))
    pcm = [(c.clamp(-1, 1) * 32767.0).to(torch.int16) for c in clips]
    want = torch.stack([O.pad_clip(torch.from_numpy((p.numpy() / 32767.0).astype(np.float32)), 100, "pad") for p in pcm])
    assert torch.equal(batch_features(pcm, 100, "pad"), want)
    with pytest.raises(NotImplementedError):
        batch_features(clips, 100, "mirror")
    with pytest.raises(AttributeError):
        batch_features([torch.zeros(101)], 100)


def test_pcm16_chunk_schedule_covers_batch():
    from audio_residual_b200.clap import CLAP_Module
    m = CLAP_Module.__new__(CLAP_Module)
    for n in (65, 256, 1000):
        b = CLAP_Module._chunk_bounds(m, n, CLAP_Module.h2d_schedule_pcm16)
        assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b, b[1:]))


class _FakeEvent:
    """CUDA-event stand-in for the planner: completed, with a given timestamp (ms)."""

    def __init__(self, t):
        self.t = t

    def query(self):
        return True

    def elapsed_time(self, other):
        return other.t - self.t


def _fake_call(sizes, c, a, b):
    rec, t = [], 0.0
    for n in sizes:
        rec.append((_FakeEvent(t), _FakeEvent(t + c * n), "copy", n))
        rec.append((_FakeEvent(t), _FakeEvent(t + a + b * n), "enc", n))
        t += 100.0
    return rec


def test_adaptive_chunk_schedule_from_measured_rates():
    """_plan_next / _pick_bounds: the fixed schedule until a call has been measured; then the candidate the two-stream pipeline
    simulation (copy(n) = c n, encode(n) = a + b n, double-buffered staging) predicts fastest. With the host to itself
    (55 GB/s: c = 0.0175 ms/clip for int16) that IS the fixed schedule; at the 23 GB/s per GPU of a saturated 8-rank host the
    chunks shrink and the tail tapers, and the prediction beats the fixed schedule by > 10 %."""
    from audio_residual_b200.clap import CLAP_Module
    m = CLAP_Module.__new__(CLAP_Module)
    m.h2d_adapt = 2                                                  # re-planning is opt-in (ARD_PIPE_ADAPT=2); 1 only measures and fits
    fixed = CLAP_Module._chunk_bounds(m, 256, CLAP_Module.h2d_schedule_pcm16)
    assert m._pick_bounds(256, torch.int16) == fixed                 # nothing measured yet
    for n in (65, 256, 300, 1000):
        for sizes in CLAP_Module._candidates(n):
            assert sum(sizes) == n and min(sizes) >= 12, (n, sizes)
    m._pipe_rates = {torch.int16: {}}
    r = m._pipe_rates[torch.int16]

    def feed(sizes, c, a, b, calls=3):       # the first two calls on a schedule (eager run, graph capture) are not fitted
        for _ in range(calls):
            r["done"] = _fake_call(sizes, c, a, b)
            m._plan_next(256, torch.int16)

    feed([32, 80, 144], 0.0175, 0.55, 0.0375, calls=2)
    assert "c" not in r
    feed([32, 80, 144], 0.0175, 0.55, 0.0375, calls=1)
    assert abs(r["c"] - 0.0175) < 1e-9 and abs(r["a"] - 0.55) < 1e-6 and abs(r["b"] - 0.0375) < 1e-8 and r["done"] is None, r
    assert m._pick_bounds(256, torch.int16) == fixed                 # host to itself: the fixed schedule stays
    t_fixed = CLAP_Module._simulate([32, 80, 144], 0.041, 0.55, 0.0375, True)
    feed([32, 80, 144], 0.041, 0.55, 0.0375, calls=1)                # the copy rate drops (saturated host): re-plan
    b = m._pick_bounds(256, torch.int16)
    sizes = [hi - lo for lo, hi in b]
    assert b[0][0] == 0 and b[-1][1] == 256 and all(x[1] == y[0] for x, y in zip(b, b[1:]))
    assert sizes[-1] <= 64 and r["predicted_ms"] < 0.9 * t_fixed, (sizes, t_fixed)
    feed(sizes, 0.041, 0.55, 0.0375, calls=3)                        # same rates measured on the new schedule: it stays
    assert m._pick_bounds(256, torch.int16) == b and r["replans"][(256, torch.int16)] == 1
    m2 = CLAP_Module.__new__(CLAP_Module)                            # default level 1: the fit is taken, the schedule stays fixed
    m2.h2d_adapt = 1
    m2._pipe_rates = {torch.int16: {}}
    for _ in range(3):
        m2._pipe_rates[torch.int16]["done"] = _fake_call([32, 80, 144], 0.041, 0.55, 0.0375)
        m2._plan_next(256, torch.int16)
    assert abs(m2._pipe_rates[torch.int16]["c"] - 0.041) < 1e-9 and m2._pick_bounds(256, torch.int16) == fixed
    # the simulation itself: copies back to back, an encode starts when its copy AND the previous encode are done
    assert abs(CLAP_Module._simulate([10, 10], 0.05, 1.0, 0.0, True) - (0.51 + 1.0 + 1.0)) < 1e-9        # copy 1 (0.51) hides under encode 0
    assert abs(CLAP_Module._simulate([10, 10], 0.5, 1.0, 0.0, True) - (2 * 5.01 + 1.0)) < 1e-9           # copy-bound: last copy + one encode
    assert m._pick_bounds(256, torch.float32) == CLAP_Module._chunk_bounds(m, 256)


def test_apply_criterion_routes_custom_losses_to_the_caller():
    from audio_residual_b200.head import apply_criterion
    z, y = torch.randn(4, 5), torch.tensor([0, 1, 2, 3])
    crit = torch.nn.CrossEntropyLoss(label_smoothing=0.1)          # not the default: the caller's own criterion runs as is
    assert torch.equal(apply_criterion(crit, z, y), crit(z, y))
    assert torch.equal(apply_criterion(lambda a, b: a.sum(), z, y), z.sum())


def test_reference_snapshot_recipe(tmp_path):
    """oracle/build_ref.py: the snapshot holds what oracle/refimport.py needs (skipped when /root/reference is not mounted)."""
    from oracle import build_ref
    if not os.path.isdir(os.path.join(build_ref.SRC, "CLAP")):
        pytest.skip("reference tree not mounted")
    assert build_ref.build(verbose=False)
    for rel in ("CLAP/src/laion_clap/clap_module/htsat.py", "CLAP/src/laion_clap/clap_module/model.py", "CLAP/src/laion_clap/training/data.py",
                "src/residual.py", "CLAP/src/laion_clap/clap_module/model_configs/HTSAT-tiny.json"):
        assert os.path.exists(os.path.join(build_ref.DST, rel)), rel
    ign = open(os.path.join(ROOT, ".gitignore")).read()
    assert "oracle/_ref/" in ign                                   # reference sources never enter the history


def test_gram_route_spectrum_equals_covariance_spectrum():
    """analyze_attention._gram_spectrum (heads with fewer samples than dimensions): the n x n Gram matrix of the centred rows has the
    non-zero eigenvalues of the D x D ddof=1 covariance; finalize_head_spectra takes that route for an accumulator that still holds
    all of its rows and the moment route otherwise, and sets the global mean either way."""
    from audio_residual_b200.analyze_attention import _gram_spectrum, _spectrum, finalize_head_spectra
    g = torch.Generator().manual_seed(4)
    n, D = 40, 96
    X = torch.rand(n, D, generator=g) ** 2
    Xd = X.double()
    ref = _spectrum(n, Xd.sum(0), Xd.t() @ Xd)
    got = _gram_spectrum(X, n, D)
    assert got.shape == (D,) and torch.allclose(got[:n - 1], ref[:n - 1], rtol=1e-9, atol=1e-14) and float(got[n:].abs().max()) == 0.0

    class Parked:                       # what finalize_head_spectra reads of a MomentAccumulator that has parked every row
        def __init__(self, rows):
            self.D, self.n, self._buf, self._fill = rows.shape[1], rows.shape[0], rows.clone(), rows.shape[0]
            self._s1 = torch.zeros(self.D, dtype=torch.float64)

        def parked_rows(self):
            return self._buf[:self._fill]

    class Moments:                      # ... and of one that has folded them
        def __init__(self, rows):
            r = rows.double()
            self.D, self.n, self.s1, self.s2 = rows.shape[1], rows.shape[0], r.sum(0), r.t() @ r

    a, b = Parked(X), Moments(X)
    wa, wb = finalize_head_spectra([a, b], n_components=16)
    assert wa.shape == (16,) and np.allclose(wa, wb, rtol=1e-9) and np.allclose(wa, ref[:16].numpy(), rtol=1e-9)
    assert torch.allclose(a.mean_global, Xd.mean(0)) and torch.allclose(b.mean_global, Xd.mean(0))
