"""A/B of the host pipeline: current audio_residual_b200/clap.py against an older copy of the file (build/ab_old/clap_old.py,
`git show <rev>:audio_residual_b200/clap.py`), same process order alternated, B = 256 int16 PCM, use_tensor=False."""
import importlib.util
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from audio_residual_b200 import weights as W  # noqa: E402
from audio_residual_b200.residual import inject_residuals  # noqa: E402
import audio_residual_b200.clap as new_clap  # noqa: E402

spec = importlib.util.spec_from_file_location("audio_residual_b200.clap_old", os.path.join(ROOT, "build", "ab_old", "clap_old.py"),
                                              submodule_search_locations=None)
old_clap = importlib.util.module_from_spec(spec)
old_clap.__package__ = "audio_residual_b200"
spec.loader.exec_module(old_clap)

B = 256
host = ((0.1 * torch.randn(B, 480000)).clamp_(-1, 1) * 32767.0).to(torch.int16).pin_memory()
mods = {}
for name, mod in (("old", old_clap), ("new", new_clap)):
    clap = mod.build_clap_module("tiny", W.make_state_dict("tiny", seed=0), device="cuda:0")
    pca, lam = W.make_pca("tiny", seed=0)
    inject_residuals(clap.model.audio_branch, pca, lam)
    mods[name] = clap
with torch.no_grad():
    for rnd in range(3):
        for name in ("old", "new"):
            clap = mods[name]
            for _ in range(3):
                clap.get_audio_embedding_from_data(host, use_tensor=False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 10
            for _ in range(n):
                clap.get_audio_embedding_from_data(host, use_tensor=False)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
            print(f"{name}: {dt * 1e3:7.2f} ms per call  {B / dt:8.0f} clips/s", flush=True)
