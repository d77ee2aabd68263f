"""e2e clips/s of CLAP_Module.get_audio_embedding_from_data for candidate host-pipeline chunk schedules (B = 256, pinned input)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import gpu_checks as G  # noqa: E402

clap, sd, _ = G.make_encoder("tiny", residual=True)
B = 256
host = (0.1 * torch.randn(B, 480000)).clamp_(-1, 1).pin_memory()
cands = [(24, 50, 80, 116, 156, 204, 256), (16, 40, 80, 136, 216, 256), (32, 60, 90, 120, 160), (20, 44, 72, 120, 160), (28, 58, 84, 100, 140),
         (24, 48, 72, 112, 150), (40, 72, 144, 200), (64, 192), (256,)]
with torch.no_grad():
    for sched in cands:
        type(clap).h2d_schedule = sched
        for _ in range(3):
            clap.get_audio_embedding_from_data(host, use_tensor=True).cpu()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 8
        for _ in range(n):
            clap.get_audio_embedding_from_data(host, use_tensor=True).cpu()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print(f"{str(clap._chunk_bounds(B)):60s} {dt * 1e3:7.2f} ms  {B / dt:8.0f} clips/s", flush=True)
