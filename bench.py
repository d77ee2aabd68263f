"""Headline benchmark: HTSAT-tiny + ResiDual (all layers) inference, clips/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

Workload (BASELINE.json configs[1]): HTSAT-tiny CLAP audio encoder with ResiDual injected in every block of all four
layers (reference-faithful doubled FFN, src/residual.py:91-96), batch 256 synthetic 10 s / 48 kHz clips per GPU,
waveform -> log-mel -> Swin encoder -> audio_projection -> L2-normalised 512-d embedding. One "step" = one batch.
N > 1: launched under torchrun, one rank per GPU, clips batch-sharded (weak scaling, no data-path collective).

JSON line keys follow the task contract: value = device-resident throughput, e2e = through the public
CLAP_Module.get_audio_embedding_from_data API with pinned-host inputs and a host read of the result, roofline = the
tcgen05 GEMM family (dominant kernel class) timed with CUDA events inside this process, cpu_baseline = the oracle port
of the reference timed on the host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "htsat_residual_clips_per_sec"
UNIT = "clips/s"
# Algorithmic FLOPs per clip, HTSAT-tiny + ResiDual on all layers, reference-faithful doubled FFN, FFT front end
# (21.10 GF excl. STFT; the ResiDual GEMM pair is folded into the out-projection here, so the GEMM
# family executes 21.10 - 1.812 = 19.29 GF/clip of tensor work; the roofline uses the flops the launches really do).
GF_PER_CLIP_TINY_FAITHFUL = 21.10


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark(self):
        """Only samples taken after this call are reported (call at the start of the timed region)."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)
        for ts, r in self.rows:
            if ts < t_mark:
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_clips_device(B, seed, device):
    """0.1*randn + three sinusoids (50..14000 Hz), clamped to [-1,1] — generated on the device."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    x = 0.1 * torch.randn(B, 480000, generator=g, device=device)
    t = torch.arange(480000, device=device, dtype=torch.float32) / 48000.0
    for _ in range(3):
        f = 50 + (14000 - 50) * torch.rand(B, 1, generator=g, device=device)
        a = 0.05 * (0.2 + 0.8 * torch.rand(B, 1, generator=g, device=device))
        x += a * torch.sin(2 * 3.14159265 * f * t[None, :])
    return x.clamp_(-1.0, 1.0)


# ------------------------------------------------------------------------------------------------- reference / CPU arm
def oracle_setup(model="tiny"):
    import torch
    from audio_residual_b200 import weights as W
    from oracle import htsat_oracle as O
    sd = W.make_state_dict(model, seed=0)
    pca, lam = W.make_pca(model, seed=0)
    ores = {l: (torch.tensor(pca[l]["mean"], dtype=torch.float32), torch.tensor(pca[l]["components"], dtype=torch.float32),
                torch.from_numpy(lam[l])) for l in pca}
    return O, sd, ores


def cpu_time_batches(nbatch, B, warm=1):
    """Oracle port of the reference path (same torch ops as the reference modules) on all host cores. Returns clips/s."""
    import torch
    O, sd, ores = oracle_setup()
    torch.set_num_threads(os.cpu_count())
    g = torch.Generator().manual_seed(99)
    wave = (0.1 * torch.randn(B, 480000, generator=g)).clamp_(-1, 1)
    times = []
    with torch.no_grad():
        for i in range(warm + nbatch):
            t0 = time.perf_counter()
            O.get_audio_embedding(wave, sd, O.CONFIGS["tiny"], ores)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    return B / (sum(times) / len(times)), sum(times) / len(times)


def run_reference(args):
    """--impl reference: the reference's own (CPU, PyTorch) implementation of the path. The reference is Python and cannot
    travel to the GPU box, so this times the oracle port (oracle/htsat_oracle.py: the same torch ops in the same order,
    pinned against the real reference by oracle/make_golden.py) on all host cores, one bounded sample of 8 clips per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    B = 8
    torch.set_num_threads(os.cpu_count())
    O, sd, ores = oracle_setup()
    g = torch.Generator().manual_seed(99)
    wave = (0.1 * torch.randn(B, 480000, generator=g)).clamp_(-1, 1)
    wl = getattr(args, "workload", "infer")
    if wl == "train":   # src/training.py:21-32 with a Linear(512,50) probe: forward + autograd backward to lambda and the classifier
        ores = {l: (mu, comp, lam.clone().requires_grad_(True)) for l, (mu, comp, lam) in ores.items()}
        torch.manual_seed(0)
        Wc, bc = (0.02 * torch.randn(50, 512)).requires_grad_(True), torch.zeros(50, requires_grad=True)
        labels = torch.randint(0, 50, (B,))

        def one():
            loss, _ = O.linear_probe_loss(wave, labels, Wc, bc, sd, O.CONFIGS["tiny"], ores)
            loss.backward()
    elif wl == "infer":
        def one():
            with torch.no_grad():
                O.get_audio_embedding(wave, sd, O.CONFIGS["tiny"], ores)
    else:
        print(json.dumps({"impl": "reference", "unavailable": f"no CPU reference arm for --workload {wl} (only infer and train are timed)"}), flush=True)
        return
    for _ in range(max(1, min(args.warmup, 2))):
        one()
    t0 = time.perf_counter()
    steps = max(1, min(args.steps, 12))
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    v = B * steps / dt
    sample = f"{steps} steps x {B} clips (bounded sample of the batch-256 workload), torch {torch.__version__} CPU fp32, {torch.get_num_threads()} threads"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(256, wl),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def config_dict(B, workload="infer"):
    common = {"batch_per_gpu": B, "clip_samples": 480000,
              "l2_policy": "per-step inputs (492 MB waveform at B=256) and activations (>2 GB) exceed the 126 MB L2",
              "weights": "random-init (seeded), random orthonormal PCA basis, lambda = 1 + 0.1 randn"}
    if workload == "train":
        return dict(common, workload="ResiDual training step (BASELINE configs[2]): HTSAT-tiny frozen, ResiDual on all layers, forward + "
                                     "backward to every lambda + 50-class Linear(512,50) probe on audio_embed, CE loss, one flat gradient "
                                     f"allreduce (27,090 floats), Adam step; batch {B} clips per GPU",
                    parallelism="batch-sharded replicas; one NCCL allreduce of the flat lambda+classifier gradient per step")
    if workload == "pca":
        return dict(common, workload="Head-representation PCA statistics (BASELINE configs[3]): HTSAT-tiny forward with capture, per-layer residual "
                                     "moments (D = 96..768) and per-(layer, head) 4096-d attention-map moments (60 heads) accumulated on the "
                                     f"tensor cores; batch {B} clips per GPU per step",
                    parallelism="clip-sharded; moments summed over ranks once at the end (outside the per-step timing)")
    if workload == "base_fusion":
        return dict(common, workload="HTSAT-base + feature fusion (aff_2d) + ResiDual on all layers, embedding throughput (BASELINE configs[4]): "
                                     f"waveform -> device get_mel x4 -> encoder -> 512-d embedding; batch {B} clips per GPU",
                    parallelism="batch-sharded replicas, no collective")
    return dict(common, workload="HTSAT-tiny + ResiDual injected in all attention blocks of layers 0-3 (reference-faithful doubled FFN), "
                                 f"inference, batch {B} clips per GPU, 10 s 48 kHz synthetic waveform -> 512-d L2-normalised embedding (BASELINE configs[1])",
                parallelism="batch-sharded replicas, no collective")


# ------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from audio_residual_b200 import lib as L
    from audio_residual_b200 import weights as W
    from audio_residual_b200.clap import build_clap_module
    from audio_residual_b200.residual import inject_residuals

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, Wm = args.batch, args.steps, max(3, args.warmup)

    wl = args.workload
    model = "base" if wl == "base_fusion" else "tiny"
    clap = build_clap_module(model, W.make_state_dict(model, seed=0), device=dev, enable_fusion=(wl == "base_fusion"))
    pca, lam = W.make_pca(model, seed=0)
    enc = clap.model.audio_branch
    wave = synth_clips_device(B, 1234 + rank, dev)
    launch_box = [0]
    if wl != "pca":
        residuals = inject_residuals(enc, pca, lam)

    if wl == "infer":
        def step():
            return enc.encode(waveform=wave, want_audio_embed=True)["audio_embed"]

        def e2e_step(host):
            return clap.get_audio_embedding_from_data(host, use_tensor=True).cpu()
        e2e_api = "CLAP_Module.get_audio_embedding_from_data(x_pinned_host, use_tensor=True).cpu()"
    elif wl == "base_fusion":
        def step():
            return enc.encode(mel_fusion=clap.fusion_mel(wave), want_audio_embed=True)["audio_embed"]

        def e2e_step(host):
            return clap.get_audio_embedding_from_data(host, use_tensor=True).cpu()
        e2e_api = "CLAP_Module(enable_fusion=True).get_audio_embedding_from_data(x_pinned_host, use_tensor=True).cpu()"
    elif wl == "train":
        from audio_residual_b200.parallel import flat_grad_allreduce
        for r in residuals.values():
            r.to(dev)
        torch.manual_seed(0)
        cls = torch.nn.Linear(512, 50).to(dev)
        params = [r.learnable for r in residuals.values()] + list(cls.parameters())
        opt = torch.optim.Adam(params, lr=1e-3)
        labels = torch.randint(0, 50, (B,), device=dev, generator=torch.Generator(device=dev).manual_seed(5 + rank))

        def train_step(x):
            opt.zero_grad(set_to_none=False)
            emb = clap.get_audio_embedding_from_data(x, use_tensor=True)
            loss = torch.nn.functional.cross_entropy(cls(emb), labels)
            loss.backward()
            flat_grad_allreduce([p.grad for p in params])
            opt.step()
            return loss.detach()

        def step():
            return train_step(wave)

        def e2e_step(host):
            return train_step(host.to(dev, non_blocking=True)).cpu()
        e2e_api = "CLAP_Module.get_audio_embedding_from_data(x.to(device), use_tensor=True) -> CE(Linear(512,50)) -> loss.backward() -> allreduce -> Adam.step(); loss.cpu()"
    else:   # pca
        from audio_residual_b200.residual import MomentAccumulator
        res_acc = [MomentAccumulator(96 << l, dev) for l in range(4)]
        heads = [4, 8, 16, 32]
        attn_acc = [[MomentAccumulator(4096, dev) for _ in range(heads[l])] for l in range(4)]

        def pca_step(x):
            out = enc.encode(waveform=x, quantize=True, want_dict=True)
            for l in range(4):
                r = out["layers_residuals"][l]
                res_acc[l].update(r.view(-1, r.shape[-1]))
                a = out["layers_attention"][l]
                a3 = a.view(a.shape[0], a.shape[1], 4096)
                for hd in range(a.shape[1]):
                    attn_acc[l][hd].update(a3[:, hd])
            return res_acc[3].s1

        def step():
            return pca_step(wave)

        def e2e_step(host):
            return pca_step(host.to(dev, non_blocking=True)).cpu()
        e2e_api = "encode(want_dict=True) + MomentAccumulator.update per layer and per (layer, head); mean vector .cpu()"

    if wl != "train":   # inference workloads run as the reference's evaluate() does (src/training.py:47): no autograd tape
        torch.set_grad_enabled(False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # started before warm-up so nvidia-smi is already streaming when the timed region begins
    for _ in range(Wm):
        out = step()
    torch.cuda.synchronize()
    L.load().ard_launch_counter_reset()
    step()
    torch.cuda.synchronize()
    launches_per_step = L.load().ard_launch_counter_read()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- timed region (device-resident inputs)
    barrier()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        out = step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ms.item()
    value = world * B * K / (ms_total / 1e3)

    # ---- e2e: public API, pinned host input, host read of the result, every step
    host = torch.empty((B, 480000), dtype=torch.float32).pin_memory()
    host.copy_(wave)
    for _ in range(2):
        e2e_step(host)
    barrier()
    t0 = time.perf_counter()
    Ke = max(2, min(K, 10))
    for _ in range(Ke):
        emb_host = e2e_step(host)
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    # bare pinned-host -> device copy rate of the same buffer: the floor the e2e step cannot beat (1.92 MB per clip over PCIe)
    dst = torch.empty_like(wave)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(3):
        dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbs = 3 * host.numel() * 4 / (time.perf_counter() - t1) / 1e9
    del dst
    e2e = {"value": world * B * Ke / te.item(), "unit": UNIT, "h2d_bytes_per_step": B * 480000 * 4, "d2h_bytes_per_step": int(emb_host.numel() * 4),
           "api": e2e_api, "steps": Ke, "h2d_gbs_measured": h2d_gbs, "h2d_bound_clips_per_s": world * h2d_gbs * 1e9 / (480000 * 4)}

    # ---- roofline of the dominant kernel class, measured live with CUDA events around every launch.
    # Every rank runs the two profiled steps (the training step holds a collective); only rank 0 records and reports.
    if rank == 0:
        L.profile_enable(True)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peaks = read_peaks()
    prof = L.profile_read()
    L.profile_enable(False)
    if world > 1:
        dist.barrier()
    tot_ms = sum(v["ms"] for v in prof.values())
    shares = {k: round(v["ms"] / tot_ms, 4) for k, v in prof.items() if v["launches"]}
    gm = prof["gemm_tc"]
    ff = prof.get("ffn_fused", {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
    achieved = gm["flops"] / (gm["ms"] * 1e-3) / 1e12
    tc_ms, tc_flops = gm["ms"] + ff["ms"], gm["flops"] + ff["flops"]
    peak = peaks["bf16_tflops_sustained"]
    roofline = {"kernel": "gemm_tc_kernel (tcgen05 GEMM family: qkv / proj+ResiDual / fc1+GELU / fc2 / merge / head)", "bound": "tensor",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                "launches_per_step": gm["launches"] // 2, "ms_per_step": gm["ms"] / 2, "flops_per_step": gm["flops"] / 2,
                "algorithmic_bytes_per_step": gm["bytes"] / 2, "achieved_hbm_gbs": gm["bytes"] / (gm["ms"] * 1e-3) / 1e9,
                "hbm_frac_of_measured": gm["bytes"] / (gm["ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "all_tcgen05_kernels": {"ms_per_step": tc_ms / 2, "tflops": tc_flops / (tc_ms * 1e-3) / 1e12,
                                        "frac": tc_flops / (tc_ms * 1e-3) / 1e12 / peak,
                                        "ffn_fused_ms_per_step": ff["ms"] / 2, "ffn_fused_launches_per_step": ff["launches"] // 2},
                "share_of_step_by_class": shares,
                "whole_step_tensor_frac": ((B * GF_PER_CLIP_TINY_FAITHFUL * 1e9) / ((ms_total / K) * 1e-3) / 1e12 / peak) if wl == "infer" else None}
    traffic_file = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(traffic_file):
        roofline["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch")

    # ---- CPU baseline (oracle port) on a bounded sample
    cpu = None
    if world == 1 and not args.no_cpu:
        v, sec = cpu_time_batches(nbatch=2, B=8, warm=1)
        import torch as _t
        cpu = {"value": v, "unit": UNIT, "cores": _t.get_num_threads(), "kind": "port",
               "sample": f"2 batches x 8 clips of the same workload ({sec:.2f} s/batch), oracle port of the reference's PyTorch CPU path, fp32"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_total / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config_dict(B, wl),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * K, "roofline": roofline, "cpu_baseline": cpu,
            "embedding_checksum": float(out.double().abs().sum().item())}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="infer", choices=["infer", "train", "pca", "base_fusion"],
                    help="infer = the headline (BASELINE configs[1]); the others measure configs[2..4] with the same JSON contract")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
