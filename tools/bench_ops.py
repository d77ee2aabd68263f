"""Op-level timings at the headline workload's shapes (HTSAT-tiny, B=256) — development tool for kernel tuning.

    python tools/bench_ops.py [gemm] [ffn] [attn] [ln] [front]

Each op is timed with CUDA events over `reps` launches after warm-up; buffers are larger than L2 where the real
workload's are. Prints time, achieved TFLOP/s and algorithmic GB/s per op.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from audio_residual_b200 import lib as L  # noqa: E402

B = int(os.environ.get("B", "256"))
dev = "cuda"


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def report(name, us, flops, bytes_):
    print(f"{name:58s} {us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s  {bytes_ / us / 1e3:8.1f} GB/s", flush=True)


def bench_gemm():
    lib = L.load()
    st = L.stream_ptr()
    for l, (T, C) in enumerate([(4096, 96), (1024, 192), (256, 384), (64, 768)]):
        M = B * T
        for name, N, K, obf, act, nres in [("qkv", 3 * C, C, 1, 0, 0), ("proj", C, C, 0, 0, 1), ("fc1+gelu", 4 * C, C, 1, 1, 0), ("fc1+geluF16", 4 * C, C, 1, 3, 0),
                                           ("fc2+2res", C, 4 * C, 0, 0, 2), ("fc1 nogelu", 4 * C, C, 1, 0, 0), ("fc2 nores", C, 4 * C, 0, 0, 0),
                                           ("fc2 bf16out", C, 4 * C, 1, 0, 0), ("fc2+1res", C, 4 * C, 0, 0, 1)]:
            A = torch.randn(M, K, device=dev).to(torch.bfloat16)
            W = torch.randn(N, K, device=dev).to(torch.bfloat16)
            bias = torch.randn(N, device=dev)
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if obf else torch.float32)
            r1 = torch.randn(M, N, device=dev) if nres >= 1 else None
            r2 = torch.randn(M, N, device=dev) if nres >= 2 else None

            def fn():
                L.check(lib.ard_gemm_bf16(L.ptr(A), K, L.ptr(W), K, L.ptr(out), N, obf, M, N, K, L.ptr(bias), act, L.ptr(r1), N, L.ptr(r2), N, st))
            us = timeit(fn)
            by = 2 * M * K + 2 * N * K + (2 if obf else 4) * M * N + 4 * M * N * nres
            report(f"stage{l} {name:11s} M={M} N={N} K={K}", us, 2.0 * M * N * K, by)
            del A, W, out, r1, r2


def bench_ffn():
    lib = L.load()
    st = L.stream_ptr()
    M, C = B * 4096, 96
    x = torch.randn(M, C, device=dev)
    r2 = torch.randn(M, C, device=dev)
    out = torch.empty_like(x)
    g, bt = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    w1 = (torch.randn(4 * C, C, device=dev) / C ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).to(torch.float16)
    b1, b2 = torch.randn(4 * C, device=dev), torch.randn(C, device=dev)
    for name, rr in (("ffn_fused_96", None), ("ffn_fused_96 +resid2", r2)):
        def fn():
            L.check(lib.ard_ffn_fused_96(L.ptr(x), L.ptr(rr), L.ptr(out), M, L.ptr(g), L.ptr(bt), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), st))
        us = timeit(fn)
        report(f"{name} M={M}", us, 2.0 * M * C * 4 * C * 2, 4.0 * M * C * (2 + (rr is not None)))
    xn = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    hb = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)

    def unfused():
        L.check(lib.ard_layernorm_bf16(L.ptr(x), L.ptr(g), L.ptr(bt), L.ptr(xn), M, C, st))
        L.check(lib.ard_gemm_bf16(L.ptr(xn), C, L.ptr(w1), C, L.ptr(hb), 4 * C, 1, M, 4 * C, C, L.ptr(b1), 3, None, 0, None, 0, st))
        L.check(lib.ard_gemm_f16(L.ptr(hb), 4 * C, L.ptr(w2), 4 * C, L.ptr(out), C, 0, M, C, 4 * C, L.ptr(b2), 0, L.ptr(x), C, None, 0, st))
    us = timeit(unfused)
    report(f"unfused LN+fc1+fc2 M={M}", us, 2.0 * M * C * 4 * C * 2, 34.0 * M * C)


def bench_ffn_wide():
    lib = L.load()
    st = L.stream_ptr()
    for T, C in ((1024, 192), (256, 384)):
        M = B * T
        x = torch.randn(M, C, device=dev)
        r2 = torch.randn(M, C, device=dev)
        out = torch.empty_like(x)
        g, bt = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        w1 = (torch.randn(4 * C, C, device=dev) / C ** 0.5).to(torch.bfloat16)
        w2 = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).to(torch.float16)
        b1, b2 = torch.randn(4 * C, device=dev), torch.randn(C, device=dev)
        b1h = 0.5 * b1
        xn = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
        hb = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
        for name, rr in ((f"ffn_fused_wide C={C}", None), (f"ffn_fused_wide C={C} +resid2", r2)):
            def fn():
                L.check(lib.ard_ffn_fused_wide(L.ptr(x), L.ptr(rr), L.ptr(out), M, C, L.ptr(g), L.ptr(bt), L.ptr(w1), L.ptr(b1h), L.ptr(w2), L.ptr(b2), st))
            us = timeit(fn)
            report(f"{name} M={M}", us, 2.0 * M * C * 4 * C * 2, 4.0 * M * C * (2 + (rr is not None)))

        def unfused():
            L.check(lib.ard_layernorm_bf16(L.ptr(x), L.ptr(g), L.ptr(bt), L.ptr(xn), M, C, st))
            L.check(lib.ard_gemm_bf16(L.ptr(xn), C, L.ptr(w1), C, L.ptr(hb), 4 * C, 1, M, 4 * C, C, L.ptr(b1), 3, None, 0, None, 0, st))
            L.check(lib.ard_gemm_f16(L.ptr(hb), 4 * C, L.ptr(w2), 4 * C, L.ptr(out), C, 0, M, C, 4 * C, L.ptr(b2), 0, L.ptr(x), C, None, 0, st))
        us = timeit(unfused)
        report(f"unfused LN+fc1+fc2 C={C} M={M}", us, 2.0 * M * C * 4 * C * 2, 34.0 * M * C)


def bench_lnqkv():
    lib = L.load()
    st = L.stream_ptr()
    M, C = B * 4096, 96
    x = torch.randn(M, C, device=dev)
    g, bt = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    w = (torch.randn(3 * C, C, device=dev) / C ** 0.5).to(torch.bfloat16)
    b = torch.randn(3 * C, device=dev)
    out = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
    xn = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    us = timeit(lambda: L.check(lib.ard_ln_qkv_96(L.ptr(x), L.ptr(g), L.ptr(bt), L.ptr(w), L.ptr(b), L.ptr(out), M, st)))
    report(f"ln_qkv_96 M={M}", us, 2.0 * M * 3 * C * C, 4.0 * M * C + 2.0 * M * 3 * C)

    def unfused():
        L.check(lib.ard_layernorm_bf16(L.ptr(x), L.ptr(g), L.ptr(bt), L.ptr(xn), M, C, st))
        L.check(lib.ard_gemm_bf16(L.ptr(xn), C, L.ptr(w), C, L.ptr(out), 3 * C, 1, M, 3 * C, C, L.ptr(b), 0, None, 0, None, 0, st))
    us = timeit(unfused)
    report(f"unfused LN + qkv M={M}", us, 2.0 * M * 3 * C * C, 8.0 * M * C + 2.0 * M * 3 * C)


def bench_attn():
    lib = L.load()
    st = L.stream_ptr()
    for l, (R, C, nH) in enumerate([(64, 96, 4), (32, 192, 8), (16, 384, 16), (8, 768, 32)]):
        M = B * R * R
        qkv = torch.randn(M, 3 * C, device=dev).to(torch.bfloat16)
        out = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
        tbl = torch.randn(225, nH, device=dev)
        for shift in (0, 4):
            def fn():
                L.check(lib.ard_window_attention(L.ptr(qkv), L.ptr(out), L.ptr(tbl), None, 1.0, 0, B, R, R, C, nH, shift, st))
            us = timeit(fn)
            report(f"stage{l} window_attention shift={shift} M={M} C={C}", us, 4.0 * M * 64 * C, 8.0 * M * C)


def bench_attn_block():
    """Window-resident attention block (attn_block.cu) vs the unfused chain it replaces (ln_qkv + window_attention + proj GEMM)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gpu_checks as G
    clap, sd, _ = G.make_encoder("tiny", residual=True)
    enc = clap.model.audio_branch
    h = enc._handle()
    lib = L.load()
    st = L.stream_ptr()
    M = B * 4096
    x = torch.randn(B, 4096, 96, device=dev)
    out = torch.empty_like(x)
    for blk in (0, 1):
        us = timeit(lambda: L.check(lib.ard_attention_block(h, 0, blk, L.ptr(x), B, L.ptr(out), st)))
        fl = M * 2.0 * (384.0 * 96 + 4.0 * (128.0 * 32 + 32.0 * 128) + 128.0 * 96)
        report(f"attn_block_96 block={blk} (shift={4 * blk}) M={M}", us, fl, 8.0 * M * 96)


def bench_stats():
    """ard_stats_accumulate_strided at the c4 shapes (one head's 4096-d attention maps out of layers_attention, B clips)."""
    from audio_residual_b200.residual import MomentAccumulator
    import torch.cuda as tc
    for l, (nW, nH) in enumerate([(64, 4), (16, 8), (4, 16), (1, 32)]):
        rows = B * nW
        a = torch.rand(rows, nH, 4096, device=dev)
        acc = MomentAccumulator(4096, a.device)
        lib = L.load()
        L.profile_enable(True)
        for _ in range(3):
            acc.update(a[:, 1])
        pr = L.profile_read()
        L.profile_enable(False)
        us = timeit(lambda: acc.update(a[:, 1]), reps=5)
        report(f"stats layer{l} rows={rows} D=4096 (gemm {pr['gemm_tc']['ms'] / 3 * 1e3:.0f} us, split+fold {pr['other']['ms'] / 3 * 1e3:.0f} us)", us,
               2.0 * 2 * rows * 4096 * 4096, 12.0 * rows * 4096 + 24.0 * 4096 * 4096)
        del a, acc
    for D, rows in ((96, B * 8192), (192, B * 2048), (384, B * 1536), (768, B * 128)):
        x = torch.rand(rows, D, device=dev)
        acc = MomentAccumulator(D, x.device)
        us = timeit(lambda: acc.update(x), reps=5)
        report(f"stats residuals rows={rows} D={D}", us, 2.0 * 2 * rows * D * D, 12.0 * rows * D)
        del x, acc


def bench_ln():
    lib = L.load()
    st = L.stream_ptr()
    for l, (T, C) in enumerate([(4096, 96), (1024, 192), (256, 384), (64, 768)]):
        M = B * T
        x = torch.randn(M, C, device=dev)
        g, bt = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        out = torch.empty(M, C, device=dev, dtype=torch.bfloat16)

        def fn():
            L.check(lib.ard_layernorm_bf16(L.ptr(x), L.ptr(g), L.ptr(bt), L.ptr(out), M, C, st))
        us = timeit(fn)
        report(f"stage{l} layernorm M={M} C={C}", us, 8.0 * M * C, 6.0 * M * C)


def bench_front():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gpu_checks as G
    clap, sd, _ = G.make_encoder("tiny")
    enc = clap.model.audio_branch
    h = enc._handle()
    lib = L.load()
    st = L.stream_ptr()
    wave = 0.1 * torch.randn(B, 480000, device=dev)
    out = torch.empty(B, 1001, 64, device=dev)

    def fn():
        L.check(lib.ard_logmel(h, L.ptr(wave), B, 480000, 0, 0, L.ptr(out), st))
    us = timeit(fn)
    report(f"stft_logmel B={B}", us, B * 501 * 5.0 * 1024 * 10, 4.0 * B * 480000 + 4.0 * B * 1001 * 64)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "ffn", "ffnw", "attn", "ln", "front"]
    print(torch.cuda.get_device_name(0), "B =", B, flush=True)
    for w in which:
        {"gemm": bench_gemm, "ffn": bench_ffn, "ffnw": bench_ffn_wide, "lnqkv": bench_lnqkv, "attn": bench_attn, "ab": bench_attn_block, "stats": bench_stats, "ln": bench_ln, "front": bench_front}[w]()
