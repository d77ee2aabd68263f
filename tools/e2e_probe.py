"""Where does the end-to-end (host-input) step spend its time? Development probe."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_residual_b200 import weights as W
from audio_residual_b200.clap import build_clap_module
from audio_residual_b200.residual import inject_residuals

dev = torch.device("cuda", 0)
torch.set_grad_enabled(False)
clap = build_clap_module("tiny", W.make_state_dict("tiny", seed=0), device=dev)
pca, lam = W.make_pca("tiny", seed=0)
inject_residuals(clap.model.audio_branch, pca, lam)
enc = clap.model.audio_branch
B = 256
host = (0.1 * torch.randn(B, 480000)).clamp_(-1, 1).pin_memory()
wave = host.to(dev)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t_cpu = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    return t_cpu * 1e3, (time.perf_counter() - t0) / reps * 1e3


for b in (64, 128, 256):
    w = wave[:b]
    cpu_ms, tot_ms = timed(lambda: enc.encode(waveform=w, want_audio_embed=True))
    print(f"encode B={b}: CPU enqueue {cpu_ms:.2f} ms, total {tot_ms:.2f} ms per call", flush=True)
for ck in (32, 64, 128, 256):
    clap.h2d_chunk = ck
    clap._copy_stream = None
    cpu_ms, tot_ms = timed(lambda: clap.get_audio_embedding_from_data(host, use_tensor=True).cpu())
    print(f"e2e pipelined chunk={ck}: {tot_ms:.2f} ms per step -> {B / tot_ms * 1e3:.0f} clips/s", flush=True)
dst = torch.empty_like(wave)
cpu_ms, tot_ms = timed(lambda: dst.copy_(host, non_blocking=True))
print(f"plain H2D 492 MB: {tot_ms:.2f} ms -> {host.numel() * 4 / tot_ms / 1e6:.1f} GB/s", flush=True)
cpu_ms, tot_ms = timed(lambda: (dst.copy_(host, non_blocking=True), enc.encode(waveform=dst, want_audio_embed=True)["audio_embed"].cpu()))
print(f"copy then encode (no overlap): {tot_ms:.2f} ms", flush=True)

# ---- timeline of one pipelined step (events on both streams, relative to a common start)
ck = 64
main = torch.cuda.current_stream()
cs = torch.cuda.Stream()
stage = [torch.empty((ck, 480000), device=dev) for _ in range(2)]
for trial in range(2):
    ev = lambda: torch.cuda.Event(enable_timing=True)
    start = ev()
    free = [ev(), ev()]
    copied = [None, None]
    marks = []
    torch.cuda.synchronize()
    start.record(main)
    cs.wait_event(start)
    for bfr in range(2):
        free[bfr].record(main)

    def start_copy(k):
        with torch.cuda.stream(cs):
            cs.wait_event(free[k % 2])
            a = ev(); a.record(cs)
            stage[k % 2].copy_(host[k * ck:(k + 1) * ck], non_blocking=True)
            b = ev(); b.record(cs)
            copied[k % 2] = b
            marks.append((f"copy{k}", a, b))
    start_copy(0)
    for k in range(4):
        if k + 1 < 4:
            start_copy(k + 1)
        main.wait_event(copied[k % 2])
        a = ev(); a.record(main)
        enc.encode(waveform=stage[k % 2], want_audio_embed=True)
        b = ev(); b.record(main)
        free[k % 2] = ev(); free[k % 2].record(main)
        marks.append((f"enc{k}", a, b))
    torch.cuda.synchronize()
    if trial == 1:
        for name, a, b in sorted(marks, key=lambda m: start.elapsed_time(m[1])):
            print(f"{name}: {start.elapsed_time(a):7.2f} -> {start.elapsed_time(b):7.2f} ms", flush=True)
