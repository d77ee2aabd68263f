"""Round-2 golden vectors from the REAL reference code (run in the build container only):

    python -m oracle.make_golden_extras     # writes tests/golden/extras_tiny_b2.npz

* head_out{l}_sample: the per-head `attn @ v` temporary of WindowAttention.forward (htsat.py:354) for every block, obtained by
  HOOKING the unmodified reference: a forward hook on each `block.attn` captures its input windows and the attention
  probabilities it returns; v is recomputed from the module's own `qkv` Linear (htsat.py:329-331) and multiplied. Compared with
  the oracle's `head_outputs` tap before storing.
* residual_module_*: src/residual.py::ResiDual.forward (:29-42) on a random [2, 64, 96] input with autograd gradients w.r.t.
  the input and `learnable` (the standalone-module contract the product's ResiDual.forward must meet).
* subset_* : forward goldens for ResiDual injected on layers (0, 2) only with n_components = 40 / 100 < D (basis rows sliced,
  src/residual.py:20-26) -- configurations the round-1 goldens did not cover.
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from audio_residual_b200 import weights as W  # noqa: E402
from oracle import htsat_oracle as O  # noqa: E402
from oracle import refimport  # noqa: E402
from oracle.make_golden import GOLDEN, golden_sample, load_into_reference, rel_err  # noqa: E402


def run(fname="extras_tiny_b2.npz", seed=0, B=2):
    ns = refimport.load()
    torch.manual_seed(0)
    clap, cfg = refimport.build_clap("tiny")
    sd = W.make_state_dict("tiny", seed=seed)
    load_into_reference(clap, sd)
    ocfg = O.CONFIGS["tiny"]
    wave = W.make_clips(B, seed=1234)
    store = {"meta_seed": np.array(seed), "meta_B": np.array(B)}
    data = [ns.get_audio_features({}, x, 480000, data_truncating="rand_trunc", data_filling="repeatpad", audio_cfg=cfg["audio_cfg"],
                                  require_grad=False) for x in wave]

    # ---- per-head attention outputs through forward hooks on the unmodified reference
    taps = {}

    def make_hook(l, b):
        def hook(mod, inputs, output):
            x = inputs[0]                                   # [B_, 64, C] windows (already rolled / partitioned)
            B_, N, C = x.shape
            qkv = mod.qkv(x).reshape(B_, N, 3, mod.num_heads, C // mod.num_heads).permute(2, 0, 3, 1, 4)   # htsat.py:329-330
            taps[(l, b)] = (output[1] @ qkv[2]).detach()    # attn [B_, nH, 64, 64] @ v [B_, nH, 64, hd]   htsat.py:354
        return hook

    handles = []
    for l, layer in enumerate(clap.audio_branch.layers):
        for b, blk in enumerate(layer.blocks):
            handles.append(blk.attn.register_forward_hook(make_hook(l, b)))
    with torch.no_grad():
        ref = clap.get_audio_output_dict(data)
        ora = O.htsat_forward({"waveform": wave}, sd, ocfg, None, head_outputs=True)
    for h in handles:
        h.remove()
    for l in range(4):
        ref_l = torch.stack([taps[(l, b)] for b in range(ocfg["depths"][l])], dim=0)
        e = rel_err(ora["head_outputs"][l], ref_l)
        assert e < 2e-5, (l, e)
        print(f"  head outputs layer {l}: {tuple(ref_l.shape)} oracle vs hooked reference rel err {e:.2e}")
        store[f"head_out{l}_sample"] = golden_sample(ref_l)
        store[f"head_out{l}_shape"] = np.array(ref_l.shape)
    assert rel_err(ora["embedding"], ref["embedding"]) < 2e-5

    # ---- standalone ResiDual module (src/residual.py:14-42) with autograd
    g = torch.Generator().manual_seed(31)
    D, K = 96, 96
    basis = torch.linalg.qr(torch.randn(D, D, generator=g, dtype=torch.float64)).Q.float()
    mean = 0.1 * torch.randn(D, generator=g)
    x = torch.randn(2, 64, D, generator=g).requires_grad_(True)
    gout = torch.randn(2, 64, D, generator=g)
    lam0 = 1 + 0.1 * torch.randn(K, generator=g)
    for tag, k in (("full", None), ("k40", 40)):
        mod = ns.residual.ResiDual(basis, mean, n_components=k)
        with torch.no_grad():
            mod.learnable.copy_(lam0[:mod.learnable.numel()])
        x.grad = None
        y = mod(x)
        y.backward(gout)
        oy = O.residual_apply(x.detach(), mean, basis[:mod.learnable.numel()], mod.learnable.detach())
        assert rel_err(oy, y.detach()) < 1e-6
        store[f"residual_module_{tag}_out"] = y.detach().numpy()
        store[f"residual_module_{tag}_dx"] = x.grad.numpy().copy()
        store[f"residual_module_{tag}_dlam"] = mod.learnable.grad.numpy().copy()
    store["residual_module_basis"] = basis.numpy()
    store["residual_module_mean"] = mean.numpy()
    store["residual_module_x"] = x.detach().numpy()
    store["residual_module_gout"] = gout.numpy()
    store["residual_module_lam"] = lam0.numpy()

    # ---- ResiDual on a subset of layers with truncated bases (n_components < D) in the FORWARD goldens
    pca, lam = W.make_pca("tiny", seed=seed)
    kcomp = {0: 40, 2: 100}
    residuals = {}
    for l, k in kcomp.items():
        small = ns.residual.ResiDual(torch.tensor(pca[l]["components"], dtype=torch.float32), torch.tensor(pca[l]["mean"], dtype=torch.float32),
                                     n_components=k)
        with torch.no_grad():
            small.learnable.copy_(torch.from_numpy(lam[l][:k]))
        residuals[l] = small
    # the reference's load_residual always builds full-rank modules (src/residual.py:161-174); a truncated module is injected the
    # way setup_residual_htsat does it (:186, :204-205): deep copy, then patch_block_with_residual on every block of the layer
    fresh = copy.deepcopy(clap.audio_branch)
    for l, r in residuals.items():
        for blk in fresh.layers[l].blocks:
            ns.residual.patch_block_with_residual(blk, r)
    saved = clap.audio_branch
    clap.audio_branch = fresh
    ores = {l: (torch.tensor(pca[l]["mean"], dtype=torch.float32), torch.tensor(pca[l]["components"][:kcomp[l]], dtype=torch.float32),
                torch.from_numpy(lam[l][:kcomp[l]].copy())) for l in kcomp}
    with torch.no_grad():
        ref2 = clap.get_audio_output_dict(data)
        ref2_emb = clap.get_audio_embedding(data)
        ora2 = O.htsat_forward({"waveform": wave}, sd, ocfg, ores)
    clap.audio_branch = saved
    for k in ("embedding", "clipwise_output"):
        assert rel_err(ora2[k], ref2[k]) < 2e-5, k
    for l in range(4):
        assert rel_err(ora2["layers_residuals"][l], ref2["layers_residuals"][l]) < 2e-5
        store[f"subset_res{l}_sample"] = golden_sample(ref2["layers_residuals"][l])
        store[f"subset_attn{l}_sample"] = golden_sample(ref2["layers_attention"][l])
    store["subset_embedding"] = ref2["embedding"].numpy()
    store["subset_audio_embed"] = ref2_emb.numpy()
    store["subset_layers"] = np.array(sorted(kcomp))
    store["subset_k"] = np.array([kcomp[l] for l in sorted(kcomp)])

    path = os.path.join(GOLDEN, fname)
    np.savez_compressed(path, **store)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    run()
