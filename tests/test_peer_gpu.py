"""Two-GPU test of the NVLink peer-memory all-reduce (ctvq_peer_*): the one-shot kernel must equal the rank-ordered
sum / world of the per-rank codebook gradients, bit-identically on both ranks, across several epochs (slot parity),
and end to end through the module's backward.  Skipped on boxes with fewer than 2 GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q, overlap=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import ct_vae_b200 as pkg
        from ct_vae_b200.dist import PeerGradComm
        C, K, D = 4, 64, 128
        comm = PeerGradComm(C * K * (D // C), dev, overlap=overlap)
        worst = 0.0
        for epoch in range(5):
            torch.manual_seed(100 * epoch + rank)
            g = torch.randn(C, K, D // C, device=dev)
            buf = comm.grad_buffer((C, K, D // C))
            buf.copy_(g)
            out = comm.allreduce_(buf)
            comm.wait()  # overlap mode: the reduced gradient is about to be read
            parts = [torch.empty_like(g) for _ in range(world)]
            dist.all_gather(parts, g)
            exp = torch.zeros_like(g)
            for p in parts:
                exp = exp + p  # rank order, like the kernel
            exp = exp * (1.0 / world)
            worst = max(worst, float((out - exp).abs().max()))
        # end to end: module backward all-reduces grad_E through the peer path
        torch.manual_seed(7)
        m = pkg.MultipleCodebookVectorQuantizer(K, D, C).to(dev)
        for qz in m.quantizers:
            qz.embedding.weight.data = torch.randn(K, D // C, device=dev) * 0.5
        pkg.attach_grad_comm(m, comm)
        torch.manual_seed(1000 + rank)
        z = torch.randn(32, D, 8, 8, device=dev, requires_grad=True)
        out, loss = m(z)
        (out.sum() * 0.01 + loss).backward()
        comm.wait()
        mine = torch.stack([qz.embedding.weight.grad for qz in m.quantizers])
        pkg.attach_grad_comm(m, None)
        for qz in m.quantizers:
            qz.embedding.weight.grad = None
        z2 = z.detach().clone().requires_grad_(True)
        out2, loss2 = m(z2)
        (out2.sum() * 0.01 + loss2).backward()
        local = torch.stack([qz.embedding.weight.grad for qz in m.quantizers])
        dist.all_reduce(local)
        local /= world
        e2e = float((mine - local).abs().max() / local.abs().max())
        torch.cuda.synchronize(dev)
        comm.close()
        q.put((rank, worst, e2e))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("overlap", [False, True])
def test_peer_allreduce_two_gpus(overlap):
    """overlap=True: the all-reduce kernel runs on a side stream (bench.py default), wait() joins it."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, overlap)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for _, worst, e2e in res:
        assert worst == 0.0, "one-shot sum must equal the rank-ordered sum exactly"
        assert e2e < 1e-5
