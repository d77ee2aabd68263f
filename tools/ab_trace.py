"""Per-role timeline of the window-resident attention block kernel's CTA 0 (development aid).

Build:  tools/build_trace.sh   (compiles attn_block.cu with -DARD_AB_TRACE into build/libard_trace.so)
Run  :  ARD_LIB_PATH=build/libard_trace.so python tools/ab_trace.py   -> cycles relative to the tile's first stamp
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("ARD_LIB_PATH", os.path.join(ROOT, "build", "libard_trace.so"))
import gpu_checks as G  # noqa: E402
from audio_residual_b200 import lib as L  # noqa: E402

B = int(os.environ.get("B", "256"))
clap, sd, _ = G.make_encoder("tiny", residual=True)
enc = clap.model.audio_branch
h = enc._handle()
lib = L.load()
lib.ard_debug_ab_trace.argtypes = [C.c_void_p]
x = torch.randn(B, 4096, 96, device="cuda")
out = torch.empty_like(x)
for _ in range(3):
    L.check(lib.ard_attention_block(h, 0, 1, L.ptr(x), B, L.ptr(out), L.stream_ptr()))
torch.cuda.synchronize()
tr = np.zeros((4, 32, 24), dtype=np.int64)
assert lib.ard_debug_ab_trace(tr.ctypes.data_as(C.c_void_p)) == 0
t0 = tr[tr > 0].min()
names = {0: "TM warp 0 (heads 0,2): 0 wait acc | 1 got | 2 drained | [3 s_full 4 S read 5 math 6 p_free 7 p_full]x2 | 13 o_full 14 ao_ready 15 x loaded 16 y_full 17 stored",
         1: "TM warp 4 (heads 1,3): same", 2: "MMA: 0 start 1 a_full 2 y_free 3 QKV issued 4 qkv_ready | [5+3h pre-wait 6+3h p_full 7+3h PV issued] | 17 ao_ready 18 proj issued",
         3: "LN warp 0: 20 loads issued 21 a_free 22 a_full arrive"}
for role in range(4):
    print(names[role])
    for i in (0, 1, 2, 10, 11):
        row = tr[role, i]
        ev = [(k, int(v - t0)) for k, v in enumerate(row) if v > 0]
        print(f"  tile {i:2d}: " + " ".join(f"{k}:{v}" for k, v in ev))
