"""Timing of the Gram-route spectrum (analyze_attention._gram_spectrum) against the D x D eigen-solve."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from audio_residual_b200.analyze_attention import _gram_spectrum, _spectrum  # noqa: E402

torch.manual_seed(0)
D = 4096
for n in (500, 1000, 2000, 2304, 3000):
    X = (torch.rand(n, D, device="cuda") ** 3)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        w = _gram_spectrum(X, n, D)
        torch.cuda.synchronize(); t1 = time.perf_counter()
    Xd = X.double()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    G = (Xd - Xd.mean(0)) @ (Xd - Xd.mean(0)).t()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    e = torch.linalg.eigvalsh(G)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"n={n}: gram route {1e3 * (t1 - t0):.1f} ms (matmul {1e3 * (t3 - t2):.1f}, eigvalsh {1e3 * (t4 - t3):.1f})", flush=True)
s2 = (X.double().t() @ X.double())
s1 = X.double().sum(0)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    w = _spectrum(3000, s1, s2)
    torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"D x D route {1e3 * (t1 - t0):.1f} ms")
