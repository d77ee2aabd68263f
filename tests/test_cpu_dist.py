"""`not gpu`: world_size-2 gloo tests of the multi-GPU host logic: clips are batch-sharded with no
data-path collective; the only collectives are the sum-allreduce of the PCA sufficient statistics and of the flat
lambda/classifier gradient buffer."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audio_residual_b200.parallel import allreduce_moments, flat_grad_allreduce, shard_range
from audio_residual_b200.residual import pca_from_moments


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((1000, 24)) @ rng.standard_normal((24, 24)) + 3.0
    lo, hi = shard_range(1000, rank, world)
    Xs = torch.from_numpy(X[lo:hi])
    n, s1, s2 = allreduce_moments(hi - lo, Xs.sum(0), Xs.T @ Xs)
    pca = pca_from_moments(n, s1.numpy(), s2.numpy())
    # gradient buffer: per-rank mean-loss grads over the local shard, combined as the global-batch mean
    g_l = torch.full((5,), float(rank + 1))
    g_w = torch.full((3, 2), float(10 * (rank + 1)))
    flat_grad_allreduce([g_l, g_w], world)
    if rank == 0:
        out.put((n, pca["explained_variance"], pca["components"], g_l.clone(), g_w.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_statistics_and_grad_allreduce_match_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    n, ev, comps, g_l, g_w = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    X = rng.standard_normal((1000, 24)) @ rng.standard_normal((24, 24)) + 3.0
    ref = pca_from_moments(1000, X.sum(0), X.T @ X)
    assert n == 1000
    assert np.allclose(ev, ref["explained_variance"], rtol=1e-10) and np.allclose(comps, ref["components"], atol=1e-8)
    assert torch.allclose(g_l, torch.full((5,), 1.5)) and torch.allclose(g_w, torch.full((3, 2), 15.0))


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 250, 2000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _spectra_worker(rank, world, port, out):
    """finalize_head_spectra (analyze_attention.py): moments reduced to round-robin owners, each rank solves its own heads."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audio_residual_b200.analyze_attention import finalize_head_spectra

    class Acc:            # a MomentAccumulator's state without the CUDA update path
        pass
    rng = np.random.default_rng(1)
    accs = []
    for h in range(5):
        X = rng.standard_normal((400, 12)) * np.linspace(2, 0.2, 12) + h
        lo, hi = shard_range(400, rank, world)
        a = Acc()
        a.D, a.n = 12, hi - lo
        a.s1 = torch.from_numpy(X[lo:hi].sum(0))
        a.s2 = torch.from_numpy(X[lo:hi].T @ X[lo:hi])
        accs.append(a)
    spectra = finalize_head_spectra(accs)
    if rank == 0:
        out.put(([a.n for a in accs], spectra))
    dist.barrier()
    dist.destroy_process_group()


def test_head_sharded_spectra_match_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_spectra_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ns, spectra = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(1)
    assert ns == [400] * 5 and len(spectra) == 5
    for h in range(5):
        X = rng.standard_normal((400, 12)) * np.linspace(2, 0.2, 12) + h
        ref = np.sort(np.linalg.eigvalsh(np.cov(X.T, ddof=1)))[::-1]
        assert np.allclose(spectra[h], ref, rtol=1e-9)


def _gram_worker(rank, world, port, out):
    """finalize_head_spectra with rank-sharded PARKED rows (fewer samples than dimensions): the rows travel, not the D x D moments."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audio_residual_b200.analyze_attention import finalize_head_spectra

    class Parked:         # what finalize_head_spectra reads of a MomentAccumulator that still holds every row
        def __init__(self, rows):
            self.D, self.n, self._buf, self._fill = rows.shape[1], rows.shape[0], rows, rows.shape[0]
            self._s1 = torch.zeros(self.D, dtype=torch.float64)

        def parked_rows(self):
            return self._buf[:self._fill]

        @property
        def s1(self):         # the real accumulator folds its parked rows when the moments are read
            return self._buf.double().sum(0)

        @property
        def s2(self):
            return self._buf.double().t() @ self._buf.double()

    class Moments:
        def __init__(self, rows):
            r = rows.double()
            self.D, self.n, self.s1, self.s2 = rows.shape[1], rows.shape[0], r.sum(0), r.t() @ r

    rng = np.random.default_rng(2)
    accs = []
    for h in range(4):
        X = torch.from_numpy((rng.random((30, 64)) ** 2 + h).astype(np.float32))
        lo, hi = (0, 11) if rank == 0 else (11, 30)                  # uneven shards
        # heads 0-2 parked on every rank (Gram route); head 3 has been folded on rank 1 (moment route for the whole head)
        accs.append(Moments(X[lo:hi]) if (h == 3 and rank == 1) else Parked(X[lo:hi]))
    spectra = finalize_head_spectra(accs)
    if rank == 0:
        out.put(([a.n for a in accs], spectra, [a.mean_global.numpy() for a in accs]))
    dist.barrier()
    dist.destroy_process_group()


def test_gram_route_with_rank_sharded_rows():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gram_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ns, spectra, means = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(2)
    assert ns == [30] * 4
    for h in range(4):
        X = (rng.random((30, 64)) ** 2 + h).astype(np.float32).astype(np.float64)
        ref = np.sort(np.linalg.eigvalsh(np.cov(X.T, ddof=1)))[::-1]
        assert np.allclose(spectra[h][:29], ref[:29], rtol=1e-8, atol=1e-12), h
        assert np.allclose(means[h], X.mean(0))


def _guard_worker(rank, world, port, out):
    """bench.RankGuard: gloo agreement between ranks for the extra records."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    g = bench.RankGuard(dist, world)
    res = [g.all_ok(True), g.all_ok(rank != 1), g.max(10.0 + rank)]
    g.barrier()
    if rank == 0:
        out.put(res)
    dist.destroy_process_group()


def test_rank_guard_agreement():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_guard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [True, False, 11.0]      # one rank's failure is seen by every rank; max over ranks
