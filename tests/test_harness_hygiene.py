"""SURVEY.md §8(f) rank 3 -- logging hygiene of the Lightning harness (experiment.py:87-110): ``fused_log_all`` logs the
same keys and values as the reference's ``log_all`` with one host synchronisation and one collective per step.
pytorch_lightning is not installed here or on the GPU box, so the experiment class is a stand-in that restates the
reference method (experiment.py:87-110) around a recording ``log_dict``."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Experiment:
    """Stand-in for VAEXperiment: `log_all` below is the reference's method restated (experiment.py:87-110)."""
    wandb_logger = False

    def __init__(self):
        self.logged, self.images = [], []

    def log_dict(self, d, sync_dist=False, batch_size=None):
        vals = dict(d)
        if sync_dist and dist.is_initialized() and dist.get_world_size() > 1:  # what Lightning does per key
            for k, v in vals.items():
                t = torch.tensor(float(v))
                dist.all_reduce(t)
                vals[k] = float(t) / dist.get_world_size()
        self.logged.append((vals, sync_dist, batch_size))

    def log_all(self, losses, batch_size, validation=False):
        if validation:
            losses = {f"val_{key}": val for key, val in losses.items()}
        to_remove = []
        for key, val in losses.items():
            if type(val) == torch.Tensor and (len(val.shape) == 0 or (len(val.shape) == 1 and val.size(0) == 1)):
                losses[key] = val.item()
            else:
                self.images.append(key)
                to_remove.append(key)
        for key in to_remove:
            del losses[key]
        self.log_dict(losses, sync_dist=True, batch_size=batch_size)


def _losses(dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {"loss": torch.rand((), generator=g).to(dev), "Reconstruction_Loss": torch.rand((), generator=g).to(dev),
            "VQ_Loss": torch.rand(1, generator=g).to(dev), "CT_Loss": torch.tensor(0.0),  # a CPU scalar, like ct_mcq_vae.py:546
            "mode": "action", "adjacency": torch.rand(4, 4, generator=g).to(dev)}


@pytest.mark.parametrize("validation", [False, True])
def test_fused_log_all_logs_what_the_reference_logs(validation):
    from ct_vae_b200 import harness
    dev = torch.device("cpu")
    ref = _Experiment()
    ref.log_all(_losses(dev), batch_size=16, validation=validation)

    class Patched(_Experiment):
        pass

    Patched.log_all = _Experiment.log_all
    assert harness.install_experiment(Patched) and not harness.install_experiment(Patched)
    exp = Patched()
    exp.log_all(_losses(dev), batch_size=16, validation=validation)
    (rv, _, rb), = ref.logged
    ours = [entry for entry in exp.logged if entry[0]]
    (ov, osync, ob), = ours
    assert set(ov) == set(rv) and ob == rb == 16
    for k in rv:
        assert ov[k] == pytest.approx(rv[k], rel=1e-7, abs=0)
    assert osync is False, "values are already reduced: no per-key collective inside log_dict"
    assert sorted(exp.images) == sorted(ref.images), "non-scalar entries keep the reference's handling"


@pytest.mark.gpu
def test_fused_log_all_synchronises_once():
    """The reference's `.item()` per scalar is one stream synchronisation each; the fused version has exactly one."""
    import warnings

    from ct_vae_b200 import harness
    dev = torch.device("cuda:0")
    losses = _losses(dev)
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("warn")
    try:
        with warnings.catch_warnings(record=True) as w_ref:
            warnings.simplefilter("always")
            _Experiment().log_all(dict(losses), batch_size=8)
        with warnings.catch_warnings(record=True) as w_ours:
            warnings.simplefilter("always")
            harness.fused_log_all(_Experiment(), dict(losses), batch_size=8)
    finally:
        torch.cuda.set_sync_debug_mode("default")
    n_ref = sum("synchroniz" in str(x.message).lower() for x in w_ref)
    n_ours = sum("synchroniz" in str(x.message).lower() for x in w_ours)
    assert n_ref >= 3 and n_ours == 1, (n_ref, n_ours)


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
        from ct_vae_b200 import harness
        calls = {"n": 0}
        real = dist.all_reduce

        def counting(*a, **k):
            calls["n"] += 1
            return real(*a, **k)

        losses = _losses(torch.device("cpu"), seed=rank)
        ref = _Experiment()
        ref.log_all(dict(losses), batch_size=4)
        exp = _Experiment()
        dist.all_reduce = counting
        try:
            harness.fused_log_all(exp, dict(losses), batch_size=4)
        finally:
            dist.all_reduce = real
        q.put((rank, ref.logged[0][0], exp.logged[-1][0], calls["n"]))
    finally:
        dist.destroy_process_group()


def test_fused_log_all_one_collective_world2():
    """sync_dist=True semantics (mean over ranks, experiment.py:110) with ONE all-reduce for all scalars of the step."""
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
    for _, ref_vals, our_vals, n_collectives in res:
        assert n_collectives == 1
        assert set(ref_vals) == set(our_vals)
        for k in ref_vals:
            assert our_vals[k] == pytest.approx(ref_vals[k], rel=1e-6)
    assert res[0][2] == res[1][2], "every rank logs the same reduced values"
