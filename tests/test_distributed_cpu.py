"""CPU, world_size 2 over gloo: the N>1 host logic - shard ranges cover the batch exactly once and the scalar
rate all-reduce reproduces the single-process aggregate (rd_loss.py:15-20 over the whole batch)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from textmae_image_compression_b200 import distributed as D


def test_shard_ranges_partition():
    for n in (1, 7, 24, 64, 288):
        for world in (1, 2, 4, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            flat = [i for b, e in spans for i in range(b, e)]
            assert flat == list(range(n))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
            rr = sorted(i for r in range(world) for i in D.shard_round_robin(n, r, world))
            assert rr == list(range(n))


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    log2lik = -torch.rand(n_total, generator=g, dtype=torch.float64) * 1e5      # per-image sum log2 likelihood
    pixels = 224.0 * 224.0
    b, e = D.shard_range(n_total, rank, world)
    local = torch.tensor([log2lik[b:e].sum().item(), pixels * (e - b)], dtype=torch.float64)
    bpp = D.aggregate_rate(local)
    per_img = D.gather_per_image((-log2lik[b:e] / pixels).float(), n_total)
    if rank == 0:
        q.put((bpp.item(), per_img.tolist(), (-log2lik.sum() / (pixels * n_total)).item(), (-log2lik / pixels).float().tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_rate_allreduce_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, per_img, want, want_per = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert abs(got - want) <= 1e-12 * abs(want)
    assert per_img == want_per


def test_aggregate_is_identity_without_group():
    t = torch.tensor([-1000.0, 50176.0], dtype=torch.float64)
    assert abs(D.aggregate_rate(t).item() - 1000.0 / 50176.0) < 1e-15
