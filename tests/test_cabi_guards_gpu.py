"""Direct C-ABI calls (ctypes, raw device pointers) with GUARD BANDS around every output buffer: whatever kernel the
dispatcher picks must not write one byte outside [B,C*d,HW] / [B,C,HW] / [C+1] / [B,Dtot,HW] / [C,K,d].
(compute-sanitizer is closed on this GPU pool, so the tests carry their own bounds check.)  Also exercises the error
convention of include/ctvq.h on a live device."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 4096  # elements on each side
SENT_F = -12345.5
SENT_I = -77


def _guarded(n, dtype, dev, sentinel):
    buf = torch.full((n + 2 * GUARD,), sentinel, dtype=dtype, device=dev)
    return buf, buf[GUARD:GUARD + n]


def _intact(buf, n, sentinel):
    return bool((buf[:GUARD] == sentinel).all()) and bool((buf[GUARD + n:] == sentinel).all())


SHAPES = [
    # B, Dtot, H, W, C, d, K, chan_stride, path
    (6, 128, 8, 8, 4, 32, 64, 1, "auto"),     # config 2: specialised tcgen05 forward + fast backward
    (700, 128, 8, 8, 4, 32, 64, 1, "auto"),   # ... large enough for the TMA-ring backward (N >= 37888)
    (3, 64, 16, 16, 1, 64, 512, 1, "auto"),   # config 1: streaming single-codebook kernel (8 units of 64 codes)
    (5, 32, 8, 8, 1, 32, 300, 1, "auto"),     # streaming kernel, d=32 (4 teams), ragged last unit, partial super-tile
    (9, 128, 8, 8, 1, 128, 200, 1, "auto"),   # streaming kernel, d=128 (2 teams, rows re-read from the slab)
    (3, 256, 4, 8, 1, 256, 100, 1, "auto"),   # streaming kernel, d=256 (1 team)
    (1200, 64, 8, 8, 1, 64, 128, 1, "auto"),  # large enough for the single-codebook shared-atomic backward
    (5, 128, 8, 8, 1, 128, 64, 1, "auto"),    # config 3
    (600, 128, 8, 8, 1, 128, 64, 1, "auto"),  # config 3, large enough for the C=1 TMA-ring backward (N >= 18944)
    (7, 64, 16, 16, 1, 64, 300, 1, "auto"),   # resident-codebook forward, two units, partial last tile
    (300, 64, 16, 16, 1, 64, 512, 1, "auto"), # config 1, large enough for the resident-accumulator ring backward (N >= 75776)
    (160, 128, 16, 16, 1, 128, 256, 1, "auto"),  # ring backward, 32-row tiles (d = 128)
    (320, 32, 16, 16, 1, 32, 256, 1, "auto"), # ring backward, 128-row tiles (d = 32)
    (150, 128, 16, 16, 4, 32, 64, 1, "auto"), # config 2 on 128x128 images: TMA-ring backward on 64-position segments
    (600, 64, 8, 8, 2, 32, 64, 1, "auto"),    # two codebooks: TMA-ring backward
    (3, 48, 8, 4, 2, 24, 50, 24, "auto"),     # generic tcgen05 kernel (K padded to 64, HW = 32)
    (2, 15, 3, 3, 5, 3, 7, 1, "auto"),        # ragged: SIMT + direct-atomic backward
    (6, 128, 8, 8, 4, 32, 64, 1, "simt"),
    (3, 64, 16, 16, 1, 64, 512, 1, "simt"),
]


@pytest.mark.parametrize("shape", SHAPES)
def test_no_write_outside_the_output_buffers(shape):
    from ct_vae_b200 import _lib
    from oracle import c_oracle as CO
    B, Dtot, H, W, C, d, K, cs, path = shape
    HW = H * W
    dev = torch.device("cuda:0")
    L = _lib.lib()
    _lib.set_path({"auto": _lib.PATH_AUTO, "simt": _lib.PATH_SIMT}[path])
    try:
        torch.manual_seed(3)
        z = torch.randn(B, Dtot, H, W, device=dev)
        books = [torch.randn(K, d, device=dev) * 0.5 for _ in range(C)]
        ptrs = (ctypes.c_void_p * C)(*[e.data_ptr() for e in books])
        ws = torch.zeros(L.ctvq_workspace_bytes(C, K, d), dtype=torch.uint8, device=dev)
        sp = torch.cuda.current_stream(dev).cuda_stream
        n_out, n_idx = B * C * d * HW, B * C * HW
        out_b, out = _guarded(n_out, torch.float32, dev, SENT_F)
        idx_b, idx = _guarded(n_idx, torch.int64, dev, SENT_I)
        loss_b, loss = _guarded(C + 1, torch.float32, dev, SENT_F)
        cnt_b, cnt = _guarded(1, torch.int64, dev, SENT_I)
        cnt.zero_()
        rc = L.ctvq_forward(z.data_ptr(), ptrs, B, Dtot, HW, C, d, K, cs, 0, 0.25, idx.data_ptr(), out.data_ptr(),
                            loss.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws.numel(), 0, sp)
        assert rc == 0, L.ctvq_strerror(rc)
        torch.cuda.synchronize()
        assert _intact(out_b, n_out, SENT_F) and _intact(idx_b, n_idx, SENT_I) and _intact(loss_b, C + 1, SENT_F)
        assert _intact(cnt_b, 1, SENT_I)
        # near-tie counter (include/ctvq.h): exact against the C oracle's count in the same evaluation order
        assert int(cnt.item()) == CO.neartie_count(z.cpu(), [e.cpu() for e in books], cs)
        assert bool((out != SENT_F).all()) and bool((idx >= 0).all()) and bool((idx < K).all())
        ref = CO.argmin(z.cpu(), [e.cpu() for e in books], cs)
        assert torch.equal(idx.cpu().view(B, C, H, W), ref)
        # backward
        g_out = torch.randn(B, C * d, H, W, device=dev)
        g_loss = torch.full((1,), 0.7, device=dev)
        n_gz, n_ge = B * Dtot * HW, C * K * d
        gz_b, gz = _guarded(n_gz, torch.float32, dev, SENT_F)
        ge_b, ge = _guarded(n_ge, torch.float32, dev, SENT_F)
        rc = L.ctvq_backward(z.data_ptr(), ptrs, idx.data_ptr(), g_out.data_ptr(), g_loss.data_ptr(), B, Dtot, HW, C, d,
                             K, cs, 0, 0.25, gz.data_ptr(), ge.data_ptr(), ws.data_ptr(), ws.numel(), 0, sp)
        assert rc == 0, L.ctvq_strerror(rc)
        torch.cuda.synchronize()
        assert _intact(gz_b, n_gz, SENT_F) and _intact(ge_b, n_ge, SENT_F)
        assert bool((gz != SENT_F).all()) and bool((ge != SENT_F).all())
        ref_gz, ref_ge = CO.backward(z.cpu(), idx.cpu().view(B, C, H, W), [e.cpu() for e in books], 0.25, g_out.cpu(),
                                     0.7, cs)
        assert float((gz.cpu().view_as(ref_gz) - ref_gz).abs().max() / ref_gz.abs().max()) < 1e-5
        assert float((ge.cpu().view_as(ref_ge) - ref_ge).abs().max() / ref_ge.abs().max().clamp_min(1e-30)) < 1e-5
    finally:
        _lib.set_path(_lib.PATH_AUTO)


def test_error_convention_on_device():
    from ct_vae_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda:0")
    z = torch.randn(2, 8, 2, 2, device=dev)
    e = torch.randn(4, 8, device=dev)
    ptrs = (ctypes.c_void_p * 1)(e.data_ptr())
    ws = torch.zeros(L.ctvq_workspace_bytes(1, 4, 8), dtype=torch.uint8, device=dev)
    idx = torch.empty(2, 1, 2, 2, dtype=torch.int64, device=dev)
    out = torch.empty_like(z)
    loss = torch.empty(2, device=dev)
    sp = torch.cuda.current_stream(dev).cuda_stream
    args = lambda dt, d, wsn: (z.data_ptr(), ptrs, 2, 8, 4, 1, d, 4, 1, dt, 0.25, idx.data_ptr(), out.data_ptr(),
                               loss.data_ptr(), None, ws.data_ptr(), wsn, 0, sp)
    assert L.ctvq_forward(*args(0, 8, ws.numel())) == 0
    assert L.ctvq_forward(*args(7, 8, ws.numel())) == -2       # unknown dtype code (CTVQ_F32 = 0 and CTVQ_BF16 = 1 exist)
    assert L.ctvq_forward(*args(0, 9, ws.numel())) == -1       # slice exceeds the channel count
    assert L.ctvq_forward(*args(0, 8, 8)) == -3                # workspace too small
    torch.cuda.synchronize()


@pytest.mark.parametrize("shape", [
    (6, 128, 8, 8, 4, 32, 64, 1),      # config 2 (bf16 kind::f16 forward / bf16 backward where specialised)
    (300, 128, 8, 8, 4, 32, 64, 1),    # ... large enough for the persistent kernels
    (3, 64, 16, 16, 1, 64, 300, 1),    # generic bf16 path (SIMT forward, tiled backward)
    (2, 15, 3, 3, 5, 3, 7, 1),         # ragged: scalar loads / stores
])
def test_bf16_through_the_c_abi(shape):
    """dtype = CTVQ_BF16: latents / output / g_out / grad_z are bf16 device buffers, codebooks stay the fp32 parameters
    (rounded to bf16 inside the kernels).  Guard bands, exact indices vs the C oracle on the rounded operands, bit-exact
    bf16 output, gradients at north_star's bf16 tolerance (2e-2; grad_E is fp32: 1e-5)."""
    from ct_vae_b200 import _lib
    from oracle import c_oracle as CO
    B, Dtot, H, W, C, d, K, cs = shape
    HW = H * W
    dev = torch.device("cuda:0")
    L = _lib.lib()
    torch.manual_seed(4)
    z = torch.randn(B, Dtot, H, W, device=dev).to(torch.bfloat16)
    books = [torch.randn(K, d, device=dev) * 0.5 for _ in range(C)]
    ptrs = (ctypes.c_void_p * C)(*[e.data_ptr() for e in books])
    ws = torch.zeros(L.ctvq_workspace_bytes(C, K, d), dtype=torch.uint8, device=dev)
    sp = torch.cuda.current_stream(dev).cuda_stream
    n_out, n_idx = B * C * d * HW, B * C * HW
    S16 = -3.0
    out_b, out = _guarded(n_out, torch.bfloat16, dev, S16)
    idx_b, idx = _guarded(n_idx, torch.int64, dev, SENT_I)
    loss_b, loss = _guarded(C + 1, torch.float32, dev, SENT_F)
    rc = L.ctvq_forward(z.data_ptr(), ptrs, B, Dtot, HW, C, d, K, cs, 1, 0.25, idx.data_ptr(), out.data_ptr(),
                        loss.data_ptr(), None, ws.data_ptr(), ws.numel(), 0, sp)
    assert rc == 0, L.ctvq_strerror(rc)
    torch.cuda.synchronize()
    assert _intact(out_b, n_out, S16) and _intact(idx_b, n_idx, SENT_I) and _intact(loss_b, C + 1, SENT_F)
    zr = z.float().cpu()
    er = [e.to(torch.bfloat16).float().cpu() for e in books]
    ref_idx = CO.argmin(zr, er, cs)
    assert torch.equal(idx.cpu().view(B, C, H, W), ref_idx)
    ref_q, ref_loss = CO.gather_st_loss(zr, ref_idx, er, 0.25, cs)
    assert torch.equal(out.float().cpu().view_as(ref_q), ref_q.to(torch.bfloat16).float())
    assert abs(float(loss[C]) - float(ref_loss[C])) <= 1e-5 * abs(float(ref_loss[C]))
    g_out = torch.randn(B, C * d, H, W, device=dev).to(torch.bfloat16)
    g_loss = torch.full((1,), 0.7, device=dev)
    n_gz, n_ge = B * Dtot * HW, C * K * d
    gz_b, gz = _guarded(n_gz, torch.bfloat16, dev, S16)
    ge_b, ge = _guarded(n_ge, torch.float32, dev, SENT_F)
    rc = L.ctvq_backward(z.data_ptr(), ptrs, idx.data_ptr(), g_out.data_ptr(), g_loss.data_ptr(), B, Dtot, HW, C, d, K, cs,
                         1, 0.25, gz.data_ptr(), ge.data_ptr(), ws.data_ptr(), ws.numel(), 0, sp)
    assert rc == 0, L.ctvq_strerror(rc)
    torch.cuda.synchronize()
    assert _intact(gz_b, n_gz, S16) and _intact(ge_b, n_ge, SENT_F)
    ref_gz, ref_ge = CO.backward(zr, ref_idx, er, 0.25, g_out.float().cpu(), 0.7, cs)
    assert float((gz.float().cpu().view_as(ref_gz) - ref_gz).abs().max() / ref_gz.abs().max()) < 2e-2
    assert float((ge.cpu().view_as(ref_ge) - ref_ge).abs().max() / ref_ge.abs().max().clamp_min(1e-30)) < 1e-5
