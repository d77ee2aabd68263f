"""Per-launch floor of each kernel family at a tiny problem size (back-to-back launches in one stream)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from audio_residual_b200 import lib as L  # noqa: E402

lib = L.load()
st = L.stream_ptr()
dev = "cuda"


def timeit(fn, reps=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for M, N, K, obf in ((128, 96, 96, 1), (128, 288, 96, 1), (4096, 288, 96, 1), (256, 1152, 384, 1), (64, 768, 3072, 0), (4096, 96, 96, 0)):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = torch.randn(N, K, device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if obf else torch.float32)
    us = timeit(lambda: L.check(lib.ard_gemm_bf16(L.ptr(A), K, L.ptr(W), K, L.ptr(out), N, obf, M, N, K, L.ptr(bias), 0, None, 0, None, 0, st)))
    print(f"gemm M={M} N={N} K={K} out16={obf}: {us:6.2f} us/launch")
for rows, C in ((128, 96), (4096, 96), (64, 768)):
    x = torch.randn(rows, C, device=dev)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    o = torch.empty(rows, C, device=dev, dtype=torch.bfloat16)
    us = timeit(lambda: L.check(lib.ard_layernorm_bf16(L.ptr(x), L.ptr(g), L.ptr(b), L.ptr(o), rows, C, st)))
    print(f"layernorm rows={rows} C={C}: {us:6.2f} us/launch")
for R, C, nH in ((64, 96, 4), (8, 768, 32)):
    qkv = torch.randn(R * R, 3 * C, device=dev).to(torch.bfloat16)
    o = torch.empty(R * R, C, device=dev, dtype=torch.bfloat16)
    tbl = torch.randn(225, nH, device=dev)
    us = timeit(lambda: L.check(lib.ard_window_attention(L.ptr(qkv), L.ptr(o), L.ptr(tbl), None, 1.0, 0, 1, R, R, C, nH, 0, st)))
    print(f"window_attention B=1 R={R} C={C}: {us:6.2f} us/launch")
x = torch.randn(4096, 96, device=dev)
o = torch.empty_like(x)
g, b = torch.ones(96, device=dev), torch.zeros(96, device=dev)
w1 = torch.randn(384, 96, device=dev).to(torch.bfloat16)
w2 = torch.randn(96, 384, device=dev).to(torch.float16)
b1, b2 = torch.randn(384, device=dev), torch.randn(96, device=dev)
us = timeit(lambda: L.check(lib.ard_ffn_fused_96(L.ptr(x), None, L.ptr(o), 4096, L.ptr(g), L.ptr(b), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), st)))
print(f"ffn_fused_96 M=4096: {us:6.2f} us/launch")
e = torch.empty(1, device=dev)
us = timeit(lambda: e.fill_(1.0))
print(f"torch fill_ (reference floor of a trivial kernel): {us:6.2f} us/launch")

# GPU-side floor: the same launches replayed from a CUDA graph (no host cost per kernel)
print("--- replayed from a CUDA graph (50 launches per graph)")
side = torch.cuda.Stream()


def graph_time(fn, n=50, reps=20):
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn(side.cuda_stream)
        torch.cuda.synchronize()
        with torch.cuda.graph(gph, stream=side):
            for _ in range(n):
                fn(side.cuda_stream)
    for _ in range(3):
        gph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / n * 1e3


for M, N, K, obf in ((128, 96, 96, 1), (4096, 288, 96, 1), (256, 1152, 384, 1), (64, 768, 3072, 0)):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = torch.randn(N, K, device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if obf else torch.float32)
    us = graph_time(lambda s: L.check(lib.ard_gemm_bf16(L.ptr(A), K, L.ptr(W), K, L.ptr(out), N, obf, M, N, K, L.ptr(bias), 0, None, 0, None, 0, s)))
    print(f"gemm M={M} N={N} K={K} out16={obf}: {us:6.2f} us/launch")
x = torch.randn(128, 96, device=dev)
g, b = torch.ones(96, device=dev), torch.zeros(96, device=dev)
o = torch.empty(128, 96, device=dev, dtype=torch.bfloat16)
print(f"layernorm rows=128 C=96: {graph_time(lambda s: L.check(lib.ard_layernorm_bf16(L.ptr(x), L.ptr(g), L.ptr(b), L.ptr(o), 128, 96, s))):6.2f} us/launch")
qkv = torch.randn(64, 3 * 768, device=dev).to(torch.bfloat16)
o2 = torch.empty(64, 768, device=dev, dtype=torch.bfloat16)
tbl = torch.randn(225, 32, device=dev)
print(f"window_attention B=1 R=8 C=768: {graph_time(lambda s: L.check(lib.ard_window_attention(L.ptr(qkv), L.ptr(o2), L.ptr(tbl), None, 1.0, 0, 1, 8, 8, 768, 32, 0, s))):6.2f} us/launch")
x4 = torch.randn(4096, 96, device=dev)
o4 = torch.empty_like(x4)
print(f"ffn_fused_96 M=4096: {graph_time(lambda s: L.check(lib.ard_ffn_fused_96(L.ptr(x4), None, L.ptr(o4), 4096, L.ptr(g), L.ptr(b), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), s))):6.2f} us/launch")
