"""Forward time vs batch size (device-resident input): the fixed per-forward cost that the chunked host pipeline pays per chunk."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import gpu_checks as G  # noqa: E402

clap, sd, _ = G.make_encoder("tiny", residual=True)
enc = clap.model.audio_branch
for B in (1, 4, 16, 40, 80, 120, 256):
    wave = (0.1 * torch.randn(B, 480000, device="cuda")).clamp_(-1, 1)
    with torch.no_grad():
        for _ in range(3):
            enc.encode(waveform=wave, want_audio_embed=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            enc.encode(waveform=wave, want_audio_embed=True)
        e1.record()
        t_issue = (time.perf_counter() - t0) / n * 1e3
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"B={B:4d}  {ms:7.3f} ms/forward  {ms / B * 1e3:8.1f} us/clip   host issue time {t_issue:6.3f} ms", flush=True)
