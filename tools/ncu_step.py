"""Minimal driver for profiler captures: build one bench.Workload, run `--warmup` + `--steps` device-resident steps and exit
(no e2e leg, no CPU baseline, no per-class timing), so `ncu` replays only the kernels of interest.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python tools/ncu_step.py --workload train
    ncu --set full --clock-control none --import-source on -k regex:gemm_dual -c 4 -o gpurun_out/dual python tools/ncu_step.py --workload train
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="infer", choices=["infer", "train", "pca", "base_fusion"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    import torch
    import bench
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    w = bench.Workload(a.workload, a.batch, dev, 0, 1)
    torch.set_grad_enabled(a.workload == "train")
    for _ in range(a.warmup):
        w.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(a.steps):
        out = w.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("checksum", float(out.double().abs().sum().item()))


if __name__ == "__main__":
    main()
