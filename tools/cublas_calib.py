"""Calibration only (not on any product path): what the vendor GEMM reaches on the same shapes, plain C = A W^T without epilogue."""
import torch
B = 256
def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for l, (T, C) in enumerate([(4096, 96), (1024, 192), (256, 384), (64, 768)]):
    M = B * T
    for name, N, K, odt in [("qkv", 3 * C, C, torch.bfloat16), ("proj", C, C, torch.float32), ("fc1", 4 * C, C, torch.bfloat16), ("fc2", C, 4 * C, torch.float32)]:
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        us = timeit(lambda: torch.matmul(A, W.t(), out=out))
        print(f"stage{l} {name:5s} M={M} N={N} K={K}  cuBLAS bf16->bf16 {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)
