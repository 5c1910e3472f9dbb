#!/usr/bin/env python
"""Headline benchmark of the quantiser hot path (driver contract; see DESIGN.md "measurement").

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels behind the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU path on host cores

A *step* is one pass of the hot path over one batch of synthetic latents: the multi-codebook quantiser of
configs/mcq_vae.yaml (C=4 codebooks, d=32, K=64 codes, encoder latents [B,128,8,8]) run forward (argmin +
gather + losses + straight-through) and backward (straight-through/commitment gradient + codebook-gradient
scatter-add), plus — for N>1 — the NCCL all-reduce of the stacked codebook gradient (the path's only
collective).  ``value`` = latent vectors (rows of the [B*H*W, 128] latent matrix, each quantised by C
codebooks) processed per second by the whole job with inputs resident in HBM; ``e2e`` = the same metric
through the public nn.Module API with the step's latents arriving from pinned HOST memory and the loss read
back to the host inside the timed region.  The same JSON line also carries the MCQ-VAE training-step
throughput (``train``: images/s, the second half of BASELINE.json's metric), the roofline of the dominant
kernel and the CPU baseline.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(C=4, K=64, D=128, H=8, W=8, beta=0.25)  # configs/mcq_vae.yaml:3-9 -> latents [B,128,8,8]
METRIC = "quantised_latents_per_sec_fwd_bwd"
UNIT = "latents/s"


_JSON_FD = None


def _quiet_stdout():
    """Route fd 1 to stderr for the whole run (NCCL / cuDNN banners are written by C code straight to fd 1) and keep
    a private duplicate for the ONE JSON line the driver parses."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes_per_row(C, d, Dtot, cs=1):
    """Bytes that MUST cross HBM per latent row: fwd reads the channels the slices touch, writes the
    quantised output and C int64 indices; bwd reads g_out, re-reads z and the indices and writes grad_z (every
    channel: untouched ones are zeros)."""
    used = min(Dtot, (C - 1) * cs + d)
    fwd = 4 * used + 4 * C * d + 8 * C
    bwd = 4 * C * d + 4 * used + 8 * C + 4 * Dtot
    return fwd, bwd


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if sm:
            hi = [s for s in sm if s >= 0.5 * max(sm)]  # samples under load
            out.update(sm_mhz=statistics.median(hi), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/): the reference's CPU arithmetic on host cores
# ------------------------------------------------------------------------------------------------------
def cpu_quantiser_step(O, z, books, g_out, beta):
    inds = O.mcq_compute_inds(z, books)
    out, loss, _ = O.mcq_compute_latents(z, inds, books, beta)
    gz, ges = O.mcq_backward(z, inds, books, beta, g_out, torch.tensor(1.0))
    return loss


def cpu_baseline(budget_s=12.0, batch=1024, max_reps=20):
    """Oracle port (same ATen CPU operators as the reference's modules) on a bounded sample of the workload."""
    from oracle import ctvq_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    C, K, D, H, W = (CFG[k] for k in "CKDHW")
    torch.manual_seed(1320)
    z = torch.randn(batch, D, H, W)
    books = [torch.randn(K, D // C) * 0.5 for _ in range(C)]
    g_out = torch.randn(batch, D, H, W)
    cpu_quantiser_step(O, z, books, g_out, CFG["beta"])  # warm-up
    times, t_all = [], time.perf_counter()
    while len(times) < max_reps and (time.perf_counter() - t_all) < budget_s:
        t0 = time.perf_counter()
        cpu_quantiser_step(O, z, books, g_out, CFG["beta"])
        times.append(time.perf_counter() - t0)
    best = min(times)
    return {"value": batch * H * W / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle/ctvq_oracle.py (reference ATen CPU ops) fwd+bwd on {batch} images = {batch * H * W} "
                      f"latents, best of {len(times)} reps", "ms_per_sample": best * 1e3}


def torch_eager_gpu_baseline(dev, batch=4096, reps=5):
    """The reference's op sequence (models/mcq_vae.py:26-64,100-127: permute, matmul distances, argmin, one-hot scatter,
    one-hot matmul, two mse_loss, straight-through, permute back; autograd backward) as STOCK torch eager on the same
    B200 — the honest GPU baseline the fused kernels replace.  Plain library calls, none of our kernels."""
    import torch.nn.functional as F
    C, K, D, H, W = (CFG[k] for k in "CKDHW")
    d = D // C
    torch.manual_seed(1320)
    books = [(torch.randn(K, d, device=dev) * 0.5).requires_grad_(True) for _ in range(C)]
    z = torch.randn(batch, D, H, W, device=dev, requires_grad=True)
    g_out = torch.randn(batch, D, H, W, device=dev)

    def step():
        outs, total = [], 0
        for i, e in enumerate(books):
            lat = z[:, i:i + d].permute(0, 2, 3, 1).contiguous()
            flat = lat.view(-1, d)
            dist = torch.sum(flat ** 2, dim=1, keepdim=True) + torch.sum(e ** 2, dim=1) - 2 * torch.matmul(flat, e.t())
            inds = torch.argmin(dist, dim=1).unsqueeze(1)
            one_hot = torch.zeros(inds.size(0), K, device=dev)
            one_hot.scatter_(1, inds, 1)
            q = torch.matmul(one_hot, e).view(lat.shape)
            total = total + F.mse_loss(q.detach(), lat) * CFG["beta"] + F.mse_loss(q, lat.detach())
            outs.append((lat + (q - lat).detach()).permute(0, 3, 1, 2).contiguous())
        out = torch.cat(outs, 1)
        torch.autograd.backward([out, total], [g_out, torch.ones((), device=dev)])
        z.grad = None
        for e in books:
            e.grad = None

    for _ in range(2):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    return {"value": batch * H * W / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch": batch,
            "what": "reference op sequence in stock torch eager (fp32, TF32 off) on this GPU"}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path, all host threads, same metric/unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ctvq_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    C, K, D, H, W = (CFG[k] for k in "CKDHW")
    batch = args.ref_batch
    torch.manual_seed(1320)
    z = torch.randn(batch, D, H, W)
    books = [torch.randn(K, D // C) * 0.5 for _ in range(C)]
    g_out = torch.randn(batch, D, H, W)
    for _ in range(max(1, args.warmup)):
        cpu_quantiser_step(O, z, books, g_out, CFG["beta"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_quantiser_step(O, z, books, g_out, CFG["beta"])
    dt = time.perf_counter() - t0
    val = batch * H * W * args.steps / dt
    sample = f"{batch} images ({batch * H * W} latents) per step: bounded sample of the workload"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(batch, 1), l2="n/a (CPU arm)"),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_config(batch_per_gpu, n_gpus):
    C, K, D, H, W = (CFG[k] for k in "CKDHW")
    return {"workload": "configs[1] MCQ-VAE quantiser (configs/mcq_vae.yaml): C=4 codebooks x K=64 codes x d=32, "
                        "latents [B,128,8,8] fp32, forward (argmin+gather+loss+straight-through) + backward "
                        "(grad_z + codebook-grad scatter-add)" + (" + NCCL all-reduce of grad_E" if n_gpus > 1 else ""),
            "batch_per_gpu": batch_per_gpu, "latents_per_gpu": batch_per_gpu * H * W, "codebooks": C,
            "num_embeddings": K, "embedding_dim": D, "chan_stride": 1,
            "parallelism": f"batch-sharded x{n_gpus}, codebooks replicated",
            "l2": "inputs larger than L2 (no flush needed)" if batch_per_gpu * D * H * W * 4 > 200e6 else "L2 flushed between steps"}


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="images per GPU per step (x64 latents each)")
    ap.add_argument("--ref-batch", type=int, default=1024)
    ap.add_argument("--train-batch", type=int, default=64, help="MCQ-VAE train-step images per GPU (configs/mcq_vae.yaml:15)")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N>1: one-shot NVLink peer-memory all-reduce of grad_E (default) or NCCL")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist

    import ct_vae_b200 as pkg
    from ct_vae_b200 import _lib
    from ct_vae_b200.dist import CodebookGradComm, PeerGradComm

    _quiet_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device: there is no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        collective = args.collective
        if collective == "peer":
            try:
                # fused: the backward kernel's last CTA runs the all-reduce (push over NVLink peer memory); everything is
                # on ONE stream in program order -- the schedule a training step can actually use (the optimizer needs the
                # reduced gradient before the next forward), nothing is hidden under the next step
                comm = PeerGradComm(CFG["C"] * CFG["K"] * (CFG["D"] // CFG["C"]), dev)
            except Exception as e:  # no CUDA IPC / peer access on this box: NCCL carries the collective instead
                print(f"[bench] peer collective unavailable ({e!r}); using NCCL", file=sys.stderr)
                collective = "nccl"
        if collective == "nccl":
            comm = CodebookGradComm(device=dev)
    _lib.lib()

    C, K, D, H, W = (CFG[k] for k in "CKDHW")
    d = D // C
    B = args.batch
    torch.manual_seed(1320 + rank)
    m = pkg.MultipleCodebookVectorQuantizer(K, D, C, CFG["beta"]).to(dev)
    torch.manual_seed(1320)
    for q in m.quantizers:
        q.embedding.weight.data = (torch.randn(K, d) * 0.5).to(dev)  # trained-like codebooks, identical on all ranks
    pkg.attach_grad_comm(m, comm)
    z = torch.randn(B, D, H, W, device=dev).requires_grad_(True)
    g_out = torch.randn(B, D, H, W, device=dev)
    g_loss = torch.ones((), device=dev)
    params = [q.embedding.weight for q in m.quantizers]

    def step(zin):
        out, loss = m(zin)
        torch.autograd.backward([out, loss], [g_out, g_loss])
        zin.grad = None
        for p in params:
            p.grad = None
        return loss

    def comm_wait():  # the overlapped all-reduce of the last step belongs to the timed region
        if comm is not None and hasattr(comm, "wait"):
            comm.wait()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step(z)
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    small = B * D * H * W * 4 <= 200e6  # inputs do not exceed L2 (126 MB) comfortably: flush it between steps
    if small:
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in evs:
            flush.zero_()
            a.record()
            step(z)
            comm_wait()
            b.record()
        sync_all()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        del flush
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step(z)
        comm_wait()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
    # per-kernel durations over the same steps (events on the launching stream = torch's current stream)
    fwd_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    bwd_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        fwd_ev[i][0].record()
        out, loss = m(z)
        fwd_ev[i][1].record()
        bwd_ev[i][0].record()
        torch.autograd.backward([out, loss], [g_out, g_loss])
        bwd_ev[i][1].record()
        z.grad = None
        for p in params:
            p.grad = None
    sync_all()
    clocks = sampler.stop() if sampler else None
    fwd_ms = statistics.mean(a.elapsed_time(b) for a, b in fwd_ev)
    bwd_ms = statistics.mean(a.elapsed_time(b) for a, b in bwd_ev)
    fwd_path = _lib.last_path()

    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    rows_per_gpu = B * H * W
    value = rows_per_gpu * world * args.steps / (ms * 1e-3)

    # ---- collective check (outside the timed region): the reduced codebook gradient against the rank-ordered sum / world
    collective_check = None
    if world > 1:
        def grads_with(c):
            pkg.attach_grad_comm(m, c)
            out, loss = m(z)
            torch.autograd.backward([out, loss], [g_out, g_loss])
            g = torch.stack([p.grad for p in params]).clone()
            z.grad = None
            for p in params:
                p.grad = None
            return g
        g_red = grads_with(comm)
        g_loc = grads_with(None)
        pkg.attach_grad_comm(m, comm)
        parts = [torch.empty_like(g_loc) for _ in range(world)]
        dist.all_gather(parts, g_loc)
        exp = torch.zeros_like(g_loc)
        for part in parts:
            exp = exp + part  # rank order, like the kernel
        exp = exp * (1.0 / world)
        reds = [torch.empty_like(g_red) for _ in range(world)]
        dist.all_gather(reds, g_red)
        identical = all(bool(torch.equal(reds[0], r_)) for r_ in reds[1:])
        rel = float((g_red - exp).abs().max() / exp.abs().max().clamp_min(1e-30))
        collective_check = {"collective": type(comm).__name__, "max_rel_err_vs_rank_ordered_sum": rel,
                            "max_abs_err": float((g_red - exp).abs().max()), "identical_on_all_ranks": identical,
                            "note": "expected value recomputed by a second local backward (fp32 atomics reorder the "
                                    "local sums, hence a relative error instead of bit identity against it)"}
        if hasattr(comm, "check"):
            comm.check()
        if rel > 1e-5 or (isinstance(comm, PeerGradComm) and not identical):
            raise RuntimeError(f"collective check failed: {collective_check}")

    # ---- e2e: public API, latents from pinned host memory each step, loss read back --------------------
    e2e_steps = max(3, min(args.steps, 10))
    host_z = torch.randn(B, D, H, W).pin_memory()
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()
    dz = torch.empty(B, D, H, W, device=dev, requires_grad=True)

    def e2e_step():
        with torch.no_grad():
            dz.copy_(host_z, non_blocking=True)
        loss = step(dz)
        comm_wait()
        host_loss.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the user reads the loss (experiment.py:96 .item())
        return float(host_loss)

    for _ in range(2):
        e2e_step()
    sync_all()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(e2e_steps):
        e2e_step()
    t1.record()
    sync_all()
    te = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = rows_per_gpu * world * e2e_steps / (float(te) * 1e-3)
    del host_z, dz

    # ---- roofline of the dominant kernel -------------------------------------------------------------------
    peak, peak_src = peaks()
    fb, bb = algorithmic_bytes_per_row(C, d, D)
    kernels = [
        {"kernel": "vq_fwd (argmin+gather+ST+loss)", "path": {1: "simt", 2: "tcgen05"}.get(fwd_path, "?"),
         "ms": fwd_ms, "alg_bytes": fb * rows_per_gpu, "gbs": fb * rows_per_gpu / (fwd_ms * 1e-3) / 1e9},
        {"kernel": "vq_backward (grad_z + codebook scatter-add)", "ms": bwd_ms, "alg_bytes": bb * rows_per_gpu,
         "gbs": bb * rows_per_gpu / (bwd_ms * 1e-3) / 1e9},
    ]
    # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu --set full capture
    # of this very workload (profiles/r1_fwd_ws_b16384.md); only quoted for the profiled batch size
    ncu_traffic = {16384: (146883328 + 511268352, 717339648 + 497255936)}.get(B)
    for i, k in enumerate(kernels):
        k["frac"] = k["gbs"] / peak
        k["traffic"] = ncu_traffic[i] if ncu_traffic else None
    dom = max(kernels, key=lambda k: k["ms"])
    roofline = {"bound": "hbm", "achieved": dom["gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                "traffic": dom["traffic"], "kernel": dom["kernel"], "peak_source": peak_src,
                "alg_bytes_per_latent": {"fwd": fb, "bwd": bb}}

    train = None
    if not args.no_train:
        train = bench_train(args, dev, world, rank, comm)
    cpu = None
    eager = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline()
        eager = torch_eager_gpu_baseline(dev)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload_config(B, world), collective=(type(comm).__name__ if comm is not None else "none")),
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": B * D * H * W * 4 * world,
                        "d2h_bytes_per_step": 4 * world, "steps": e2e_steps},
                "gpu_launches": (2 + (1 if world > 1 and not getattr(comm, "fuses_backward", False) else 0)) * args.steps,
                "collective_check": collective_check, "roofline": roofline, "kernels": kernels, "clocks": clocks, "cpu_baseline": cpu, "gpu_eager_baseline": eager, "train": train}
        emit(line)
    if world > 1:
        if comm is not None:
            comm.close()
        dist.destroy_process_group()


def bench_train(args, dev, world, rank, comm):
    """Second half of BASELINE.json's metric: MCQ-VAE (configs/mcq_vae.yaml) training images/s — conv
    encoder/decoder on stock cuDNN exactly like the reference, the quantiser on our kernels, Adam lr 5e-4
    (mcq_vae.yaml:23); N>1: torch DDP for the model, batch-sharded."""
    import torch.distributed as dist

    import ct_vae_b200 as pkg
    from ct_vae_b200.harness import GraphedTrainer, MCQVAEShell

    B = args.train_batch
    torch.manual_seed(1320)
    model = MCQVAEShell(3, 128, 64, [64, 128, 256], 0.25, 64, 4).to(dev)
    trainer = GraphedTrainer(model, (B, 3, 64, 64), dev, lr=5e-4, world=world)
    torch.manual_seed(1320 + rank)
    x = torch.rand(B, 3, 64, 64, device=dev)  # Shapes3D images are in [0,1] after ToTensor (dataset.py:72-75)
    steps, warm = max(20, args.steps), max(5, args.warmup)
    for _ in range(warm):
        trainer.step(x)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        trainer.step(x)
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    # e2e: images from pinned host memory each step + loss read back to the host (experiment.py:96 .item())
    hx = torch.rand(B, 3, 64, 64).pin_memory()
    for _ in range(2):
        float(trainer.step(hx))
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        float(trainer.step(hx))
    torch.cuda.synchronize(dev)
    te = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    return {"metric": "mcq_vae_train_images_per_sec", "value": B * world * steps / (ms * 1e-3), "unit": "images/s",
            "batch_per_gpu": B, "steps": steps, "ms_per_step": ms / steps,
            "e2e": {"value": B * world * steps / (float(te) * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": B * 3 * 64 * 64 * 4 * world, "d2h_bytes_per_step": 4 * world},
            "cuda_graph": trainer.graph is not None,
            "model": "MCQVAEShell = layer structure of models/mcq_vae.py:142-317 (10.1 M params), cuDNN convs, "
                     "ctvq quantiser, Adam lr 5e-4; step replayed from two CUDA graphs (fwd+bwd, Adam) with one eager NCCL all-reduce of the flat gradient between them for N>1"}


if __name__ == "__main__":
    main()
