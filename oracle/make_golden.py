"""Generate tests/golden/*.npz from the REAL reference code (run in the build container only).

    python -m oracle.make_golden            # writes tests/golden/htsat_tiny_b2.npz, htsat_base_fusion_b2.npz, ...

For every case the reference modules (imported via oracle/refimport.py, unmodified) and the oracle restatement
(oracle/htsat_oracle.py) are run on identical seeded inputs and weights (audio_residual_b200/weights.py); the script
asserts they agree to float32 round-off and stores the REFERENCE outputs. Large tensors (layers_residuals,
layers_attention, framewise outputs) are stored as fixed strided samples plus float64 checksums to keep the
fixtures small; `golden_sample()` below defines the sampling and is what the tests use on the candidate side.
"""
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from audio_residual_b200 import weights as W  # noqa: E402
from oracle import htsat_oracle as O  # noqa: E402
from oracle import refimport  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_sample(t, n=4096):
    """Deterministic strided sample of a tensor (flattened), used for the big outputs."""
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    while step > 1 and (step % 2 == 0 or step % 3 == 0):   # never alias with the channel count (96*2^l / 128*2^l)
        step -= 1
    return f[::step][:n].to(torch.float32).cpu().numpy()


def checksum(t):
    f = t.detach().to(torch.float64)
    return np.array([f.sum().item(), f.abs().sum().item(), (f * f).sum().item()])


def load_into_reference(clap, sd):
    ab = clap.audio_branch
    own = ab.state_dict()
    missing = [k for k in own if k not in sd and not k.endswith(("relative_position_index", "attn_mask",
                                                                 "num_batches_tracked")) and not k.startswith("head.")
               and "mel_conv2d" not in k and "fusion_model" not in k]
    assert not missing, missing
    ab.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
    clap.audio_projection.load_state_dict({"0.weight": sd["audio_projection.0.weight"], "0.bias": sd["audio_projection.0.bias"],
                                           "2.weight": sd["audio_projection.2.weight"], "2.bias": sd["audio_projection.2.bias"]})
    clap.eval()


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def pack_outputs(prefix, out, store):
    store[prefix + "embedding"] = out["embedding"].detach().float().numpy()
    store[prefix + "clipwise_output"] = out["clipwise_output"].detach().float().numpy()
    store[prefix + "framewise_sample"] = golden_sample(out["framewise_output"])
    store[prefix + "fine_sample"] = golden_sample(out["fine_grained_embedding"])
    for l in range(4):
        store[prefix + f"res{l}_sample"] = golden_sample(out["layers_residuals"][l])
        store[prefix + f"res{l}_cks"] = checksum(out["layers_residuals"][l])
        store[prefix + f"attn{l}_sample"] = golden_sample(out["layers_attention"][l])
        store[prefix + f"attn{l}_cks"] = checksum(out["layers_attention"][l])


def compare_dicts(tag, ref, ora, tol):
    worst = 0.0
    for k in ("embedding", "clipwise_output", "framewise_output", "fine_grained_embedding"):
        e = rel_err(ora[k], ref[k]); worst = max(worst, e)
        assert e < tol, (tag, k, e)
    for l in range(4):
        for k in ("layers_residuals", "layers_attention"):
            e = rel_err(ora[k][l], ref[k][l]); worst = max(worst, e)
            assert e < tol, (tag, k, l, e)
    print(f"  [{tag}] oracle vs reference: worst rel err {worst:.2e}")


def run_case(model_name, fusion, B, seed, fname):
    ns = refimport.load()
    torch.manual_seed(0)
    clap, cfg = refimport.build_clap(model_name, enable_fusion=fusion, fusion_type="aff_2d" if fusion else "None")
    sd = W.make_state_dict(model_name, seed=seed)
    load_into_reference(clap, sd)
    ocfg = O.CONFIGS[model_name]
    wave = W.make_clips(B, seed=1234)
    store = {"meta_model": np.array(model_name), "meta_seed": np.array(seed), "meta_B": np.array(B),
             "meta_fusion": np.array(int(fusion))}
    audio_cfg = cfg["audio_cfg"]

    def featurise(w):   # hook.py:175-188 with use_tensor=True (no quantise)
        return [ns.get_audio_features({}, x, 480000, data_truncating="fusion" if fusion else "rand_trunc",
                                      data_filling="repeatpad", audio_cfg=audio_cfg, require_grad=False) for x in w]

    # ---------------- plain encoder (config 1: forward + capture) ----------------
    with torch.no_grad():
        data = featurise(wave)
        ref = clap.get_audio_output_dict(data)
        ref_emb = clap.get_audio_embedding(data)
        if fusion:
            inp = {"mel_fusion": torch.stack([d["mel_fusion"] for d in data])}
            store["mel_fusion_sample"] = golden_sample(inp["mel_fusion"])
        else:
            inp = {"waveform": wave}
        ora = O.htsat_forward(inp, sd, ocfg, None, enable_fusion=fusion)
        ora_emb = O.audio_projection(ora["embedding"], sd)
    compare_dicts("plain", ref, ora, 2e-5)
    assert rel_err(ora_emb, ref_emb) < 2e-5
    pack_outputs("plain_", ref, store)
    store["plain_audio_embed"] = ref_emb.numpy()

    if not fusion:
        # front-end intermediates (pin K1-K4)
        with torch.no_grad():
            ab = clap.audio_branch
            spec = ab.spectrogram_extractor(wave)
            lm = ab.logmel_extractor(spec)
            xb = ab.bn0(lm.transpose(1, 3)).transpose(1, 3)
            img = ab.reshape_wav2img(xb)
            pe = ab.patch_embed(img)
            o_lm = O.logmel(O.stft_power(wave, sd), sd)
            o_img = O.reshape_wav2img(O.bn0_eval(o_lm, sd))
            o_pe = O.patch_embed(o_img, sd)
        assert rel_err(o_lm, lm) < 1e-5 and rel_err(o_img, img) < 1e-5 and rel_err(o_pe, pe) < 1e-5
        store["logmel_sample"] = golden_sample(lm)
        store["logmel_cks"] = checksum(lm)
        store["img_sample"] = golden_sample(img)
        store["patch_embed_sample"] = golden_sample(pe)
        # quantised-input route (evaluation path, hook.py:177 / src/residual.py:210)
        q_ref = torch.from_numpy(ns.int16_to_float32(ns.float32_to_int16(wave.numpy())))
        assert torch.equal(q_ref, torch.from_numpy(O.int16_roundtrip_np(wave.numpy())))
        assert torch.equal(ns.residual.quantize_tensor(wave), O.quantize_tensor(wave))
        store["quant_sample"] = golden_sample(q_ref, 1024)

    # ---------------- ResiDual on all layers (config 2) + training-step gradients (config 3) ----------------
    pca, lam = W.make_pca(model_name, seed=seed)
    tmp = tempfile.mkdtemp()
    files = {}
    for l, d in pca.items():
        files[l] = os.path.join(tmp, f"layer_{l}")
        D = d["components"].shape[0]
        with open(files[l], "wb") as f:
            pickle.dump({"components": d["components"], "mean": d["mean"], "explained_variance": np.ones(D),
                         "explained_variance_ratio": np.ones(D) / D, "n_components": D, "input_dim": D,
                         "num_samples": 1}, f)
    new_htsat, residuals = ns.residual.setup_residual_htsat(clap.audio_branch, files, [0, 1, 2, 3])
    clap.audio_branch = new_htsat
    for l, r in residuals.items():
        r.learnable.data = torch.from_numpy(lam[l]).clone()
    ores = {l: (torch.tensor(pca[l]["mean"], dtype=torch.float32), torch.tensor(pca[l]["components"], dtype=torch.float32),
                torch.from_numpy(lam[l]).clone().requires_grad_(True)) for l in pca}
    with torch.no_grad():
        ref2 = clap.get_audio_output_dict(data)
        ref2_emb = clap.get_audio_embedding(data)
        ora2 = O.htsat_forward(inp, sd, ocfg, ores, enable_fusion=fusion)
        ora2_emb = O.audio_projection(ora2["embedding"], sd)
    compare_dicts("residual", ref2, ora2, 2e-5)
    assert rel_err(ora2_emb, ref2_emb) < 2e-5
    pack_outputs("residual_", ref2, store)
    store["residual_audio_embed"] = ref2_emb.numpy()

    text = W.make_text_embeds(50, 512, seed=7)
    labels = torch.from_numpy(np.random.default_rng(11).integers(0, 50, size=B))
    store["labels"] = labels.numpy()
    # reference training step: src/training.py:21-32 (encoder stays in eval mode, hook.py:173)
    emb = clap.get_audio_embedding(featurise(wave))
    sims = emb.float() @ text.T
    loss = torch.nn.CrossEntropyLoss()(sims, labels)
    loss.backward()
    oloss, osims = O.zero_shot_loss(wave, labels, text, sd, ocfg, ores) if not fusion else (None, None)
    store["train_loss"] = np.array(loss.item())
    store["train_sims"] = sims.detach().numpy()
    for l, r in residuals.items():
        store[f"lambda_grad{l}"] = r.learnable.grad.numpy().copy()
    if not fusion:
        oloss.backward()
        assert abs(oloss.item() - loss.item()) < 1e-5
        for l in residuals:
            e = rel_err(ores[l][2].grad, residuals[l].learnable.grad)
            assert e < 5e-4, (l, e)
            print(f"  lambda-grad layer {l}: oracle vs reference rel err {e:.2e}")
    # linear-probe head gradients (src/linear.py:23-45): frozen encoder, trainable Linear(512, 50)
    Wc = (W._randn(seed, "cls.weight", 50, 512) * np.float32(np.sqrt(2.0 / 512))).astype(np.float32)
    Wc_t = torch.from_numpy(Wc).requires_grad_(True)
    bc_t = torch.zeros(50, requires_grad=True)
    logits = torch.nn.functional.linear(emb.detach(), Wc_t, bc_t)
    l2 = torch.nn.functional.cross_entropy(logits, labels)
    l2.backward()
    store["cls_weight"] = Wc
    store["cls_loss"] = np.array(l2.item())
    store["cls_logits"] = logits.detach().numpy()
    store["cls_weight_grad"] = Wc_t.grad.numpy()
    store["cls_bias_grad"] = bc_t.grad.numpy()

    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, fname)
    np.savez_compressed(path, **store)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def run_pca_case(fname="pca_moments.npz"):
    """IncrementalPCA (as driven by compute_pca_components, src/residual.py:110,137-138) vs the moment form."""
    from sklearn.decomposition import IncrementalPCA
    rng = np.random.default_rng(5)
    D, n_batches, per = 96, 10, 512
    A = rng.standard_normal((D, D)) * np.linspace(2.0, 0.1, D)[None, :]
    mu = rng.standard_normal(D)
    ipca = IncrementalPCA(n_components=None)
    s1 = np.zeros(D); s2 = np.zeros((D, D)); n = 0
    for i in range(n_batches):
        X = (rng.standard_normal((per, D)) @ A.T + mu).astype(np.float32)
        ipca.partial_fit(X)
        Xd = X.astype(np.float64)
        s1 += Xd.sum(0); s2 += Xd.T @ Xd; n += per
    got = O.pca_from_moments(n, s1, s2)
    assert np.allclose(got["mean"], ipca.mean_, atol=1e-6)   # sklearn keeps float32 inputs in float32
    assert np.allclose(got["explained_variance"], ipca.explained_variance_, rtol=2e-5)
    assert np.allclose(got["components"], ipca.components_, atol=5e-6), "components incl. sign convention"
    path = os.path.join(GOLDEN, fname)
    np.savez_compressed(path, n=n, s1=s1, s2=s2, mean=ipca.mean_, components=ipca.components_,
                        explained_variance=ipca.explained_variance_, explained_variance_ratio=ipca.explained_variance_ratio_)
    print(f"wrote {path}")


def run_keys_case(fname="reference_state_dict.json"):
    """Key names + shapes of the reference's audio_branch / audio_projection state_dict (drop-in contract for the
    product's module tree and for ard_set_weight)."""
    import json
    out = {}
    for name, fusion in (("tiny", False), ("base", True)):
        clap, _ = refimport.build_clap(name, enable_fusion=fusion, fusion_type="aff_2d" if fusion else "None")
        d = {k: list(v.shape) for k, v in clap.audio_branch.state_dict().items()}
        d.update({"audio_projection." + k: list(v.shape) for k, v in clap.audio_projection.state_dict().items()})
        out[f"{name}{'_fusion' if fusion else ''}"] = d
    path = os.path.join(GOLDEN, fname)
    json.dump(out, open(path, "w"), indent=0, sort_keys=True)
    print(f"wrote {path}")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    run_keys_case()
    run_pca_case()
    run_case("tiny", False, 2, 0, "htsat_tiny_b2.npz")
    run_case("base", True, 2, 1, "htsat_base_fusion_b2.npz")
    run_case("base", False, 2, 2, "htsat_base_b2.npz")          # HTSAT-base on the waveform route (no feature fusion)
