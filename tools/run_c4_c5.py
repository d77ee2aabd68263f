"""BASELINE configs[3] and [4] on one GPU (development / profiles script).

c4: head-representation PCA over 2000 synthetic clips, all layers: per-layer residual moments (compute_pca_components' statistics)
    and per-(layer, head) 4096-d attention-map moments (run_PCA's statistics), then the eigen-solves.
c5: HTSAT-base + fusion + ResiDual embedding throughput sweep, batch 64..4096, through bench.py --workload base_fusion.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def c5(batches):
    out = []
    for b in batches:
        steps = max(2, min(10, 2048 // b))
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "base_fusion", "--batch", str(b), "--steps", str(steps),
                            "--no-cpu"], capture_output=True, text=True, timeout=900)
        line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else ""
        try:
            d = json.loads(line)
            out.append({"batch": b, "clips_per_s": d["value"], "ms_per_step": d["ms_per_step"], "e2e_clips_per_s": d["e2e"]["value"],
                        "gemm_tflops": d["roofline"]["achieved"], "clocks": d["clocks"]})
        except Exception as e:  # noqa: BLE001
            out.append({"batch": b, "error": (p.stderr or str(e))[-400:]})
        print(json.dumps(out[-1]), flush=True)
    return out


def c4(n_clips=2000, batch=125):
    import torch
    from audio_residual_b200 import weights as W
    from audio_residual_b200.clap import build_clap_module
    from audio_residual_b200.analyze_attention import HeadPCA
    from audio_residual_b200.residual import MomentAccumulator
    sys.path.insert(0, ROOT)
    import bench
    dev = torch.device("cuda", 0)
    torch.set_grad_enabled(False)
    clap = build_clap_module("tiny", W.make_state_dict("tiny", seed=0), device=dev)
    enc = clap.model.audio_branch
    heads = [4, 8, 16, 32]
    res_acc = [MomentAccumulator(96 << l, dev) for l in range(4)]
    head_pca = [[HeadPCA(4096, dev) for _ in range(heads[l])] for l in range(4)]
    wave = bench.synth_clips_device(batch, 1234, dev)
    enc.encode(waveform=wave, quantize=True, want_dict=True)       # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for step in range(n_clips // batch):
        wave = bench.synth_clips_device(batch, 5000 + step, dev)
        out = enc.encode(waveform=wave, quantize=True, want_dict=True)
        for l in range(4):
            r = out["layers_residuals"][l]
            res_acc[l].update(r.view(-1, r.shape[-1]))
            a = out["layers_attention"][l]
            a3 = a.view(a.shape[0], a.shape[1], 4096)
            for hd in range(a.shape[1]):
                head_pca[l][hd].partial_fit(a3[:, hd])
    torch.cuda.synchronize()
    t_stats = time.perf_counter() - t0
    t1 = time.perf_counter()
    comps = [acc.pca() for acc in res_acc]
    t_res = time.perf_counter() - t1
    t2 = time.perf_counter()
    for l in range(4):
        for hd in range(heads[l]):
            head_pca[l][hd].finalize(None)
    torch.cuda.synchronize()
    t_heads = time.perf_counter() - t2
    ev0 = head_pca[0][0].explained_variance_
    pr = float(ev0.sum() ** 2 / (ev0 ** 2).sum())
    res = {"clips": n_clips, "batch": batch, "forward_plus_moments_s": t_stats, "clips_per_s": n_clips / t_stats,
           "residual_pca_eigh_s (4 layers, numpy float64 on host)": t_res, "head_spectrum_eigvalsh_s (60 x 4096^2, float64 on GPU)": t_heads,
           "samples_per_head_layer0": head_pca[0][0].acc.n, "participation_ratio_layer0_head0": pr,
           "residual_components_shapes": [list(c["components"].shape) for c in comps]}
    print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    result = {}
    if what in ("c4", "all"):
        result["c4"] = c4()
    if what in ("c5", "all"):
        result["c5"] = c5([64, 128, 256, 512, 1024, 2048, 4096])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(result, open(os.path.join(ROOT, "gpurun_out", f"c4_c5_{what}.json"), "w"), indent=1)
