"""GPU parity of the STREAMING single-codebook tcgen05 kernel (ctvq_tc_stream.cu), forced through
ctvq_set_path(CTVQ_PATH_TC_STREAM): codebooks of any size stream through a TMA ring, D = 32 / 64 / 128 / 256.
Bar (north_star): indices equal the C oracle's on EVERY row (same arithmetic contract), outputs bit-exact given the
indices, loss within 1e-5 relative.  Also pins torch.argmin's non-finite rule (models/vq_vae.py:35: the first NaN wins)
on every kernel path.
"""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def env():
    import ct_vae_b200 as pkg
    from ct_vae_b200 import _lib
    from oracle import c_oracle, ctvq_oracle
    assert torch.cuda.is_available()
    _lib.lib()
    return pkg, _lib, ctvq_oracle, c_oracle


@pytest.fixture(autouse=True)
def _reset_path():
    yield
    from ct_vae_b200 import _lib
    _lib.set_path(_lib.PATH_AUTO)


def _module(pkg, K, D, kind, dev):
    m = pkg.VectorQuantizerMS(K, D, 0.25)
    if kind == "trained":
        m.embedding.weight.data = torch.randn(K, D) * 0.5
    return m.to(dev)


@pytest.mark.parametrize("cfg", [
    # (name, B, D, H, W, K, codebook): ragged unit counts, partial super-tiles, every D variant
    ("d32_k100_partial", 3, 32, 8, 8, 100, "trained"),
    ("d32_k256_init", 40, 32, 16, 16, 256, "init"),          # 20 super-tiles of 512 rows, tie-heavy
    ("d32_k4096", 2, 32, 16, 16, 4096, "trained"),           # 64 units, one super-tile
    ("d64_k512_init", 40, 64, 8, 8, 512, "init"),             # configs/vq_vae.yaml codebook at HW=64: 5 super-tiles
    ("d64_k1_single_code", 5, 64, 8, 8, 1, "trained"),
    ("d64_k65", 7, 64, 8, 8, 65, "trained"),
    ("d64_k1000", 9, 64, 16, 16, 1000, "trained"),
    ("d128_k64_cfg3", 33, 128, 8, 8, 64, "trained"),          # configs/ct_mcq_vae.yaml quantiser: 8.25 super-tiles of 256
    ("d128_k1024", 8, 128, 16, 16, 1024, "trained"),
    ("d256_k300", 9, 256, 4, 8, 300, "init"),                 # 2.25 super-tiles of 128 rows
    ("d256_k2048", 2, 256, 16, 16, 2048, "trained"),
    # the largest codebooks of BASELINE.json configs[3] (K = 16384), every row checked against the C oracle
    ("d32_k16384", 8, 32, 16, 16, 16384, "trained"),          # 256 units, 4 super-tiles of 512 rows
    ("d32_k16384_init", 4, 32, 16, 16, 16384, "init"),        # tie-heavy: |e| ~ 1/K, the window holds many codes
    ("d256_k16384", 4, 256, 16, 16, 16384, "trained"),        # 8 super-tiles of 128 rows, 8 blocks per unit
])
def test_stream_vs_oracles(env, cfg):
    pkg, _lib, O, CO = env
    name, B, D, H, W, K, kind = cfg
    dev = torch.device("cuda:0")
    torch.manual_seed(4321)
    m = _module(pkg, K, D, kind, dev)
    z_cpu = torch.randn(B, D, H, W)
    z = z_cpu.to(dev)
    _lib.set_path(_lib.PATH_TC_STREAM)
    with torch.no_grad():
        out, loss, inds = m(z, inds=True)
        only_inds = m.compute_inds(z)
    torch.cuda.synchronize()
    book = [m.embedding.weight.detach().cpu()]
    inds_cpu = inds.cpu().reshape(B, 1, H, W)
    assert torch.equal(inds_cpu, CO.argmin(z_cpu, book)), "indices differ from the C oracle"
    assert torch.equal(only_inds.cpu().reshape(B, 1, H, W), inds_cpu), "argmin-only launch differs from the fused launch"
    assert m.near_tie_rows() == 2 * CO.neartie_count(z_cpu, book), "near-tie counter (two launches) differs from the C oracle"
    ref_out, ref_loss, _ = O.mcq_compute_latents(z_cpu, inds_cpu, book, 0.25)
    assert torch.equal(out.cpu(), ref_out)
    assert rel_err(loss.cpu(), ref_loss) < TOL
    # reference ATen arithmetic: only counted near-ties may differ
    ref_inds = O.mcq_compute_inds(z_cpu, book)
    near, hard = O.classify_index_mismatches(z_cpu, book[0], inds_cpu[:, 0], ref_inds[:, 0])
    assert hard == 0
    print(f"{name}: rows={B * H * W} near-tie mismatches vs reference arithmetic: {near}")


def test_stream_pair_batching(env):
    """CT pair (models/ct_mcq_vae.py:530,536): x and y in one launch == two launches, on the streaming kernel."""
    pkg, _lib, O, CO = env
    dev = torch.device("cuda:0")
    torch.manual_seed(7)
    m = pkg.MultipleCodebookVectorQuantizer(64, 128, 1, 0.25)   # the C=1 quantiser of configs/ct_mcq_vae.yaml
    m.quantizers[0].embedding.weight.data = torch.randn(64, 128) * 0.5
    m = m.to(dev)
    x = torch.randn(16, 128, 8, 8, device=dev)
    y = torch.roll(x, 3, 0) * 1.25
    _lib.set_path(_lib.PATH_TC_STREAM)
    ix, iy = m.compute_inds_pair(x, y)
    assert torch.equal(ix, m.compute_inds(x)) and torch.equal(iy, m.compute_inds(y))
    book = [m.quantizers[0].embedding.weight.detach().cpu()]
    assert torch.equal(ix.cpu().reshape(16, 1, 8, 8), CO.argmin(x.cpu(), book))
    assert torch.equal(iy.cpu().reshape(16, 1, 8, 8), CO.argmin(y.cpu(), book))


# (the streaming kernel is single-codebook: that combination is not generated)
@pytest.mark.parametrize("path,shape", [(p, s) for s in [(64, 1, 512), (128, 4, 64), (32, 1, 256)]
                                        for p in ["simt", "tc", "stream"] if not (p == "stream" and s[1] != 1)])
def test_non_finite_rows_follow_torch_argmin(env, path, shape):
    """NaN / inf latents: torch.argmin returns the first NaN distance (verified against the live reference when the
    goldens were minted); every kernel path and the C oracle must agree with torch on those rows."""
    pkg, _lib, O, CO = env
    D, C, K = shape
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    d = D // C
    if C == 1:
        m = pkg.VectorQuantizerMS(K, D, 0.25)
        books = [m.embedding.weight]
    else:
        m = pkg.MultipleCodebookVectorQuantizer(K, D, C, 0.25)
        books = [q.embedding.weight for q in m.quantizers]
    for e in books:
        e.data = torch.randn(K, d) * 0.5
    m = m.to(dev)
    z_cpu = torch.randn(4, D, 8, 8)
    z_cpu[0, 3, 1, 2] = float("nan")
    z_cpu[1, 0, 0, 0] = float("inf")
    z_cpu[2, d - 1, 7, 7] = float("-inf")
    z_cpu[3, 5, 4, 4] = 3.0e38   # |z|^2 overflows to +inf
    _lib.set_path({"simt": _lib.PATH_SIMT, "tc": _lib.PATH_TC, "stream": _lib.PATH_TC_STREAM}[path])
    inds = m.compute_inds(z_cpu.to(dev))  # every generated (path, shape) pair is inside that path's domain
    books_cpu = [e.detach().cpu() for e in books]
    ref = O.mcq_compute_inds(z_cpu, books_cpu)   # the reference's ATen arithmetic, torch.argmin semantics
    got = inds.cpu().reshape(ref.shape)
    assert torch.equal(got, CO.argmin(z_cpu, books_cpu).reshape(ref.shape)), "kernel and C oracle disagree"
    bad = ~torch.isfinite(z_cpu).all(dim=1) | (z_cpu.abs() > 1e19).any(dim=1)   # rows with a non-finite distance
    for c in range(C):
        assert torch.equal(got[:, c][bad], ref[:, c][bad]), "non-finite rows must follow torch.argmin"


@pytest.mark.parametrize("cfg", [
    # (name, B, D, H, W, K): batches large enough for the single-codebook backward kernel (ctvq_bwd_c1.cu) to be chosen
    ("bwd_d32_k256", 1024, 32, 16, 16, 256),
    ("bwd_d64_k512_cfg1_codebook", 1024, 64, 16, 16, 512),     # configs/vq_vae.yaml at B=1024
    ("bwd_d128_k256_tm64", 2048, 128, 8, 8, 256),              # accumulator leaves room for 64-row tiles only
    ("bwd_d64_k200_ragged", 515, 64, 8, 8, 200),               # last tile partial, K not a power of two
    ("bwd_d64_k1024_global_acc", 128, 64, 16, 16, 1024),       # [K,d] exceeds shared memory: coalesced atomics to grad_E
    ("bwd_d256_k256_global_acc", 130, 256, 16, 16, 256),
])
def test_single_codebook_backward_at_scale(env, cfg):
    """grad_z and the codebook-gradient scatter-add against the C oracle at sizes where the
    shared-atomic single-codebook kernel runs; 1e-5 relative (atomics reorder the fp32 sums)."""
    pkg, _lib, O, CO = env
    name, B, D, H, W, K = cfg
    dev = torch.device("cuda:0")
    torch.manual_seed(99)
    m = _module(pkg, K, D, "trained", dev)
    z_cpu = torch.randn(B, D, H, W)
    g_out = torch.randn(B, D, H, W)
    z = z_cpu.to(dev).requires_grad_(True)
    out, loss, inds = m(z, inds=True)
    (out * g_out.to(dev)).sum().add(0.7 * loss).backward()
    torch.cuda.synchronize()
    book = [m.embedding.weight.detach().cpu()]
    inds_cpu = inds.cpu().reshape(B, 1, H, W)
    gz, ge = CO.backward(z_cpu, inds_cpu, book, 0.25, g_out, 0.7)
    assert rel_err(z.grad.cpu(), gz) < TOL
    assert rel_err(m.embedding.weight.grad.cpu(), ge.reshape(K, D)) < TOL


def test_tf32_operands_are_truncated(env):
    """The candidate window of every tf32 kernel is CENTRED on the mean truncation loss (ctvq_common.cuh, kTruncC /
    kTf32Eps): that is only rigorous if tcgen05.mma kind::tf32 truncates its fp32 operands (toward zero) to 10 explicit
    mantissa bits.  Pin it through the raw-TMEM dump of the generic kernel: a part that rounded to nearest instead would
    return 1 + 2^-10 for the operands below and this test -- not a rare index mismatch -- would say so."""
    import ctypes
    pkg, _lib, O, CO = env
    dev = torch.device("cuda:0")
    B, D, H, W, K = 2, 32, 8, 8, 64
    m = pkg.VectorQuantizerMS(K, D).to(dev)
    E = torch.zeros(K, D, device=dev)
    E[:, 0] = 1.0
    E[1, 0] = 1.0 + 2.0 ** -11 + 2.0 ** -20          # B operand just above the rounding midpoint
    m.embedding.weight.data = E
    z = torch.zeros(B, D, H, W, device=dev)
    vals = [1.0 + 2.0 ** -11 + 2.0 ** -20, 1.0 + 2.0 ** -10 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11, -(1.0 + 2.0 ** -11 + 2.0 ** -20),
            1.0 + 2.0 ** -10 - 2.0 ** -23]
    for i, v in enumerate(vals):
        z[0, 0, 0, i] = v
    z[0, 0, 1, 0] = 1.0
    dump = torch.full((128, K), float("nan"), device=dev)
    L = _lib.lib()
    L.ctvq_debug_set_tc_dump(ctypes.c_void_p(dump.data_ptr()))
    try:
        _lib.set_path(_lib.PATH_TC)
        m(z, inds=True)
        torch.cuda.synchronize()
    finally:
        L.ctvq_debug_set_tc_dump(None)
    assert not torch.isnan(dump[: len(vals), 0]).any(), "the generic tcgen05 kernel did not run (dump untouched)"
    for i, v in enumerate(vals):
        trunc = float(torch.tensor(v).view(torch.int32).bitwise_and(~0x1FFF).view(torch.float32))
        assert float(dump[i, 0]) == trunc, f"A operand {v!r}: tensor core used {float(dump[i, 0])!r}, truncation gives {trunc!r}"
    assert float(dump[8, 1]) == 1.0, f"B operand 1 + 2^-11 + 2^-20 was not truncated to 1.0: {float(dump[8, 1])!r}"
