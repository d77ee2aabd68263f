"""Per-role timeline of the fused FFN kernel's CTA 0 (development aid).

Build:  tools/build_trace.sh   (compiles the fused FFN kernels with -DARD_FFN_TRACE into build/libard_trace.so)
Run  :  python tools/ffn_trace.py [--r2]      -> cycles relative to the first stamp, one line per tile / chunk
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "build", "libard_trace.so"))
vp, ll = C.c_void_p, C.c_longlong
lib.ard_ffn_fused_96.argtypes = [vp, vp, vp, ll, vp, vp, vp, vp, vp, vp, vp]
lib.ard_debug_ffn_trace.argtypes = [vp]

M, Cc = 256 * 4096, 96
dev = "cuda"
x = torch.randn(M, Cc, device=dev)
r2 = torch.randn(M, Cc, device=dev) if "--r2" in sys.argv else None
out = torch.empty_like(x)
g, bt = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
w1 = (torch.randn(4 * Cc, Cc, device=dev) / Cc ** 0.5).to(torch.bfloat16)
w2 = (torch.randn(Cc, 4 * Cc, device=dev) / (4 * Cc) ** 0.5).to(torch.float16)
b1, b2 = torch.randn(4 * Cc, device=dev), torch.randn(Cc, device=dev)
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
for _ in range(3):
    rc = lib.ard_ffn_fused_96(p(x), p(r2), p(out), M, p(g), p(bt), p(w1), p(b1), p(w2), p(b2), None)
    assert rc == 0, rc
torch.cuda.synchronize()
tr = np.zeros((8, 64, 8), dtype=np.int64)
assert lib.ard_debug_ffn_trace(tr.ctypes.data_as(C.c_void_p)) == 0
t0 = tr[tr > 0].min()
rel = np.where(tr > 0, tr - t0, -1)
names = {0: "epilogue  [wait y_full | got | tile done]", 1: "fc1 issue [start | a1_full | h_free | committed]",
         2: "fc2 issue [start | a2_full | y_free | committed]", 3: "LN        [loads issued | a1_free | a1_full arrive]",
         4: "GELU g0   [start | h_full | math done | a2_free | a2_full arrive]", 5: "GELU g1   [same]"}
for role in range(6):
    print(names[role])
    n = 8 if role in (0, 3) else 48
    for i in range(n):
        row = rel[role, i]
        if (row >= 0).any():
            tag = f"tile {i}" if role in (0, 3) else f"tile {i // 6} chunk {i % 6}"
            print(f"  {tag:16s}", " ".join(f"{v:8d}" for v in row[:5] if v >= 0))
