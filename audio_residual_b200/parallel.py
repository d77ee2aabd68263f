"""Multi-GPU plumbing for the path: one process per GPU, clips batch-sharded, full weight replica per
GPU, NO data-path collective for inference. The only exchanges are two latency-bound sum-allreduces over NCCL/NVSwitch:
the PCA sufficient statistics {n, sum x, sum x x^T} at the end of a pass, and one flat fp32 buffer holding the lambda and
classifier gradients per training step. Device-agnostic (gloo on CPU in the tests, nccl on the GPUs)."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) slice of n units (clips) for this rank."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def allreduce_moments(n, s1, s2):
    """Sum {n, s1[D], s2[D,D]} (float64) over ranks in one flat buffer; returns (n_total, s1, s2)."""
    if not _active():
        return int(n), s1, s2
    D = s1.numel()
    flat = torch.cat([torch.tensor([float(n)], dtype=torch.float64, device=s1.device), s1.reshape(-1).double(), s2.reshape(-1).double()])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return int(round(flat[0].item())), flat[1:1 + D].reshape(s1.shape), flat[1 + D:].reshape(s2.shape)


def flat_grad_allreduce(grads, world=None):
    """Average per-rank gradients (each computed with a mean loss over the local shard) in ONE allreduce of a flat fp32
    buffer, written back in place. For HTSAT-tiny with ResiDual on all layers + a 50-class probe this is 27,090 floats."""
    if not _active():
        return grads
    world = world or dist.get_world_size()
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= world
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].reshape(g.shape))
        off += g.numel()
    return grads
