"""GPU tests of the NCCL wrapper `ctvq_allreduce_codebook_grad` (include/ctvq.h) -- the collective BASELINE.json's
north_star names: "the codebook gradient is all-reduced with NCCL over NVLink; that is the only collective".  It replaces
the share of Lightning's DDP bucket all-reduce that carries vq_layer.*.embedding.weight.grad (run.py:99).

  * world = 1 on ONE GPU (never skipped): the C ABI is driven directly -- dlopen NCCL, unique id, communicator, the
    pre-mul-sum all-reduce on a non-default stream -- and must return scale * x.
  * world = 2 (skipped below 2 GPUs): CodebookGradComm end to end against the rank-ordered sum / world.
"""
import ctypes
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def test_allreduce_codebook_grad_world1_through_the_c_abi():
    from ct_vae_b200 import _lib
    from ct_vae_b200.dist import _find_libnccl
    L = _lib.lib()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    assert L.ctvq_nccl_load((_find_libnccl() or "libnccl.so.2").encode()) == 0
    ident = ctypes.create_string_buffer(128)
    assert L.ctvq_nccl_unique_id(ident) == 0
    comm = ctypes.c_void_p()
    assert L.ctvq_nccl_comm_init(ctypes.byref(comm), 1, 0, ident.raw, 0) == 0
    try:
        torch.manual_seed(0)
        C, K, d = 4, 64, 32   # the stacked [C,K,d] gradient of configs/mcq_vae.yaml: 32 KB
        x = torch.randn(C, K, d, device=dev)
        for scale in (1.0, 0.5, 0.125):
            g = x.clone()
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                rc = L.ctvq_allreduce_codebook_grad(comm, g.data_ptr(), g.numel(), scale, 0, s.cuda_stream)
            assert rc == 0, L.ctvq_strerror(rc)
            s.synchronize()
            assert torch.equal(g, x * scale), "one rank: sum over ranks, times scale (exact: scale is a power of two)"
        # bad arguments are refused, a zero-length message is a no-op
        assert L.ctvq_allreduce_codebook_grad(None, x.data_ptr(), x.numel(), 1.0, 0, None) == -4
        assert L.ctvq_allreduce_codebook_grad(comm, x.data_ptr(), 0, 1.0, 0, None) == 0
    finally:
        torch.cuda.synchronize(dev)
        assert L.ctvq_nccl_comm_destroy(comm) == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import ct_vae_b200 as pkg
        from ct_vae_b200.dist import CodebookGradComm
        comm = CodebookGradComm(device=dev)
        worst = 0.0
        for it in range(3):
            torch.manual_seed(10 * it + rank)
            g = torch.randn(4, 64, 32, device=dev)
            parts = [torch.empty_like(g) for _ in range(world)]
            dist.all_gather(parts, g)
            exp = torch.zeros_like(g)
            for p in parts:
                exp = exp + p
            exp = exp / world
            out = comm.allreduce_(g.clone())
            worst = max(worst, float((out - exp).abs().max() / exp.abs().max()))
        # end to end through the module's backward
        torch.manual_seed(7)
        m = pkg.MultipleCodebookVectorQuantizer(64, 128, 4).to(dev)
        pkg.attach_grad_comm(m, comm)
        torch.manual_seed(1000 + rank)
        z = torch.randn(32, 128, 8, 8, device=dev, requires_grad=True)
        out, loss = m(z)
        (out.sum() * 0.01 + loss).backward()
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
def test_codebook_grad_comm_two_gpus():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for _, worst, e2e in res:
        assert worst < 1e-6 and e2e < 1e-5
