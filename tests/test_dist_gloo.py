"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: batch sharding covers the batch exactly
once, and CodebookGradComm reproduces DDP's gradient averaging (sum / world) for the stacked [C,K,d] grad."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ct_vae_b200.dist import CodebookGradComm, shard_batch
        from oracle import ctvq_oracle as O
        torch.manual_seed(0)
        C, K, d, B = 4, 16, 8, 10
        z = torch.randn(B, C - 1 + d, 3, 3)
        books = [torch.randn(K, d) for _ in range(C)]
        g_out = torch.randn(B, C * d, 3, 3)
        lo, hi = shard_batch(B, world, rank)
        # per-rank partial codebook gradient of the per-rank MEAN loss (what each DDP replica computes)
        inds = O.mcq_compute_inds(z[lo:hi], books)
        _, ges = O.mcq_backward(z[lo:hi], inds, books, 0.25, g_out[lo:hi], torch.tensor(1.0))
        ge = torch.stack(ges)
        comm = CodebookGradComm(device=None)
        comm.allreduce_(ge)
        # expected: average over ranks of the per-rank gradients
        exp = torch.zeros_like(ge)
        for r in range(world):
            a, b = shard_batch(B, world, r)
            ir = O.mcq_compute_inds(z[a:b], books)
            _, gr = O.mcq_backward(z[a:b], ir, books, 0.25, g_out[a:b], torch.tensor(1.0))
            exp += torch.stack(gr) / world
        q.put((rank, float((ge - exp).abs().max()), (lo, hi), comm.launches))
    finally:
        dist.destroy_process_group()


def test_codebook_grad_allreduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][2] == (0, 5) and res[1][2] == (5, 10)
    for _, err, _, launches in res:
        assert err < 1e-6
        assert launches == 1


def test_shard_batch_partitions_exactly():
    from ct_vae_b200.dist import shard_batch
    for total in (1, 7, 64, 65, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_batch(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
