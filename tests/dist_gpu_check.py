"""Run under torchrun on N >= 2 GPUs (tests/test_gpu_dist.py spawns it): SURVEY §4 item (5), N-GPU sharded result == 1-GPU result.

  c3: one ResiDual training step on a global batch of 2N clips, clips batch-sharded, mean loss per shard, ONE flat allreduce of the
      lambda + classifier gradients (parallel.flat_grad_allreduce)  ==  the same step on one GPU over all 2N clips.
  c4: residual moments of a layer + 4096-d attention-map moments of two heads over the sharded clips, summed with
      MomentAccumulator.allreduce / finalize_head_spectra  ==  the single-GPU statistics.
Rank 0 also computes the single-GPU reference and writes the comparison as one JSON line.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def train_step(clap, residuals, cls, wave, labels, allreduce):
    from audio_residual_b200.head import cross_entropy, head_logits
    from audio_residual_b200.parallel import flat_grad_allreduce
    params = [r.learnable for r in residuals.values()] + list(cls.parameters())
    for p in params:
        p.grad = None
    emb = clap.get_audio_embedding_from_data(wave, use_tensor=True)
    loss = cross_entropy(head_logits(emb, cls.weight, cls.bias), labels)
    loss.backward()
    grads = [p.grad.to(wave.device) for p in params]
    if allreduce:
        flat_grad_allreduce(grads)
    return loss.detach(), grads


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import gpu_checks as G
    from audio_residual_b200.analyze_attention import finalize_head_spectra
    from audio_residual_b200.parallel import shard_range
    from audio_residual_b200.residual import MomentAccumulator, inject_residuals
    W = G.W
    n_clips = 2 * world
    wave_all = W.make_clips(n_clips, seed=555)
    labels_all = torch.arange(n_clips) % 50
    lo, hi = shard_range(n_clips, rank, world)

    def build():
        from audio_residual_b200.clap import build_clap_module
        clap = build_clap_module("tiny", W.make_state_dict("tiny", seed=0), device=dev)
        pca, lam = W.make_pca("tiny", seed=0)
        res = inject_residuals(clap.model.audio_branch, pca, lam)
        torch.manual_seed(0)
        cls = torch.nn.Linear(512, 50).to(dev)
        return clap, res, cls

    out = {"world": world}
    # ---- c3: sharded step
    clap, res, cls = build()
    loss, grads = train_step(clap, res, cls, wave_all[lo:hi].to(dev), labels_all[lo:hi].to(dev), allreduce=True)
    loss_sum = loss.clone() * (hi - lo)
    dist.all_reduce(loss_sum)
    # ---- c4: sharded statistics (plain encoder, capture)
    with torch.no_grad():
        from audio_residual_b200.clap import build_clap_module
        clap2 = build_clap_module("tiny", W.make_state_dict("tiny", seed=0), device=dev)      # plain encoder on THIS rank's GPU
        o = clap2.model.audio_branch.encode(waveform=wave_all[lo:hi].to(dev), quantize=True, want_dict=True)
        racc = MomentAccumulator(192, dev)
        racc.update(o["layers_residuals"][1])
        a2 = o["layers_attention"][2]
        heads = [MomentAccumulator(4096, dev) for _ in range(3)]
        for i, h in enumerate((0, 7, 15)):
            heads[i].update(a2[:, h].reshape(a2.shape[0], 4096))
        racc.allreduce()
        spectra = finalize_head_spectra(heads, n_components=8)
    if rank == 0:
        # ---- single-GPU reference over all clips
        clap_s, res_s, cls_s = build()
        loss_s, grads_s = train_step(clap_s, res_s, cls_s, wave_all.to(dev), labels_all.to(dev), allreduce=False)
        out["loss_abs"] = abs(loss_sum.item() / n_clips - loss_s.item())
        out["grad_rel"] = [rel(a, b) for a, b in zip(grads, grads_s)]
        with torch.no_grad():
            o = clap2.model.audio_branch.encode(waveform=wave_all.to(dev), quantize=True, want_dict=True)
            r1 = MomentAccumulator(192, dev)
            r1.update(o["layers_residuals"][1])
            a2 = o["layers_attention"][2]
            h1 = [MomentAccumulator(4096, dev) for _ in range(3)]
            for i, h in enumerate((0, 7, 15)):
                h1[i].update(a2[:, h].reshape(a2.shape[0], 4096))
            torch.cuda.synchronize()
        out["moments_n"] = [racc.n, r1.n]
        out["moments_s1_rel"] = rel(racc.s1, r1.s1)
        out["moments_s2_rel"] = rel(racc.s2, r1.s2)
    # every rank takes part in the single-GPU spectra call? no: it has no collective when run by rank 0 alone with world > 1,
    # so compute the reference spectra from the moments directly
    if rank == 0:
        from audio_residual_b200.analyze_attention import _spectrum
        ref_sp = [_spectrum(a.n, a.s1, a.s2)[:8].cpu().numpy() for a in h1]
        out["spectra_rel"] = [float(abs(s - r).max() / abs(r).max()) for s, r in zip(spectra, ref_sp)]
        print("DIST_CHECK " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
