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
PCM = "--pcm16" in sys.argv     # int16 PCM host batch through use_tensor=False (the evaluation route bench.py's e2e leg times)
host = (0.1 * torch.randn(B, 480000)).clamp_(-1, 1)
host = ((host * 32767.0).to(torch.int16) if PCM else host).pin_memory()
call = (lambda: clap.get_audio_embedding_from_data(host, use_tensor=False)) if PCM else \
    (lambda: clap.get_audio_embedding_from_data(host, use_tensor=True).cpu())
cands = [(32, 80, 144, 256), (32, 110, 256), (16, 72, 256), (24, 96, 256), (20, 84, 256), (16, 64, 176), (12, 48, 96, 256), (24, 232), (48, 208),
         (256,)] if PCM else [(24, 50, 80, 116, 156, 204, 256), (16, 40, 80, 136, 216, 256), (32, 60, 90, 120, 160), (20, 44, 72, 120, 160), (28, 58, 84, 100, 140),
         (24, 48, 72, 112, 150), (40, 72, 144, 200), (64, 192), (256,)]
with torch.no_grad():
    for sched in cands:
        if PCM:
            type(clap).h2d_schedule_pcm16 = sched
        else:
            type(clap).h2d_schedule = sched
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 8
        for _ in range(n):
            call()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print(f"{str(clap._chunk_bounds(B, sched)):60s} {dt * 1e3:7.2f} ms  {B / dt:8.0f} clips/s", flush=True)
