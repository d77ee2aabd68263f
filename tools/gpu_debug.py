"""Run every GPU parity check and print all metrics (does not stop at the first failure). Usage on the GPU box:
    python tools/gpu_debug.py [ops|encoder|all]
"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import gpu_checks as G  # noqa: E402


def run(name, fn, *a, **k):
    t0 = time.time()
    try:
        r = fn(*a, **k)
        torch.cuda.synchronize()
        print(f"[ok ] {name}: {r}  ({time.time() - t0:.2f}s)", flush=True)
    except Exception as e:
        print(f"[ERR] {name}: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()
        try:
            torch.cuda.synchronize()
        except Exception as e2:
            print("CUDA context is broken:", e2, flush=True)
            sys.exit(3)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), torch.version.cuda, flush=True)
    if what == "pair":
        for (M, N, K) in [(2048, 2304, 768), (5000, 1536, 384), (65536, 384, 1536), (2304, 128, 384), (4097, 192, 3072), (300 * 128, 1152, 384)]:
            run(f"gemm pair f32+res M{M} N{N} K{K}", G.check_gemm, M, N, K, False, nres=1)
            run(f"gemm pair bf16 M{M} N{N} K{K}", G.check_gemm, M, N, K, True)
    if what in ("ops", "all"):
        for (M, N, K) in [(128, 96, 64), (128, 96, 96), (256, 128, 128), (1000, 288, 96), (4096, 384, 96), (512, 96, 384),
                          (300, 768, 768), (2048, 2304, 768), (640, 527, 4608), (8192, 192, 384), (4096, 256, 1024)]:
            run(f"gemm f32 M{M} N{N} K{K}", G.check_gemm, M, N, K, False)
            run(f"gemm bf16 M{M} N{N} K{K}", G.check_gemm, M, N, K, True)
        for (M, N, K) in [(2048, 2304, 768), (5000, 1536, 384), (65536, 384, 1536), (2304, 128, 384), (4097, 192, 3072), (300 * 128, 1152, 384)]:
            run(f"gemm pair f32+res M{M} N{N} K{K}", G.check_gemm, M, N, K, False, nres=1)
            run(f"gemm pair bf16 M{M} N{N} K{K}", G.check_gemm, M, N, K, True)
        run("gemm gelu bf16", G.check_gemm, 4096, 384, 96, True, act=1)
        run("gemm relu f32", G.check_gemm, 512, 512, 768, False, act=2)
        run("gemm 2 resid f32", G.check_gemm, 4096, 96, 384, False, nres=2)
        run("gemm 1 resid nobias f32", G.check_gemm, 1024, 192, 384, False, bias=False, nres=1)
        for Cd in (96, 128, 192, 384, 768, 1536):
            run(f"layernorm C{Cd}", G.check_layernorm, 1000, Cd)
        for (B, R, Cd, nH, sh) in [(2, 64, 96, 4, 0), (2, 64, 96, 4, 4), (2, 32, 192, 8, 4), (3, 16, 384, 16, 4), (2, 8, 768, 32, 4),
                                   (2, 64, 128, 4, 4), (2, 16, 512, 16, 0)]:
            run(f"window_attention B{B} R{R} C{Cd} nH{nH} shift{sh}", G.check_window_attention, B, R, Cd, nH, sh)
        run("logmel", G.check_logmel)
    if what in ("stats", "all"):
        for (rows, D, st, calls) in [(1000, 96, False, 1), (5000, 192, False, 2), (777, 768, False, 1), (3000, 4096, True, 2), (130, 384, True, 1)]:
            run(f"stats rows{rows} D{D} strided={st}", G.check_stats, rows, D, st, calls=calls)
        run("pca moments layer 0 vs oracle", G.check_pca_moments_vs_oracle, 0, 2)
        run("pca moments layer 3 vs oracle", G.check_pca_moments_vs_oracle, 3, 2)
    if what in ("bwd", "all"):
        for Cd in (96, 128, 192, 384, 768, 1536):
            run(f"layernorm_bwd C{Cd}", G.check_layernorm_bwd, 1000, Cd)
        run("layernorm_bwd no add", G.check_layernorm_bwd, 777, 384, with_add=False)
        for (B, R, Cd, nH, sh) in [(2, 64, 96, 4, 0), (2, 64, 96, 4, 4), (2, 32, 192, 8, 4), (3, 16, 384, 16, 4), (2, 8, 768, 32, 4),
                                   (2, 64, 128, 4, 4), (2, 16, 512, 16, 0)]:
            run(f"window_attention_bwd B{B} R{R} C{Cd} nH{nH} shift{sh}", G.check_window_attention_bwd, B, R, Cd, nH, sh)
        run("training step tiny vs golden (reference loss.backward)", G.check_training_step_vs_golden)
        run("training step tiny layers (1,) vs oracle", G.check_training_step_vs_oracle, "tiny", 2, (1,))
        run("training step tiny layers (2,3) vs oracle", G.check_training_step_vs_oracle, "tiny", 3, (2, 3))
    if what in ("encoder", "all"):
        run("encoder tiny plain vs oracle", G.check_encoder_vs_oracle, "tiny", 2, False)
        run("encoder tiny residual vs oracle", G.check_encoder_vs_oracle, "tiny", 2, True)
        run("encoder tiny vs golden", G.check_encoder_vs_golden, "htsat_tiny_b2.npz")
        run("encoder base fusion vs golden", G.check_encoder_vs_golden, "htsat_base_fusion_b2.npz")
        run("fusion featuriser + base from waveform", G.check_fusion_featuriser)


if __name__ == "__main__":
    main()
