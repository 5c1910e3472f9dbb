"""The tf32 candidate window is CENTRED on the mean operand-truncation loss (ct_vae_b200/csrc/ctvq_common.cuh: kTruncC,
kTf32Eps).  This CPU test restates the bound the kernels rely on and checks it numerically:

    tcgen05.mma.kind::tf32 truncates fp32 operands to 10 explicit mantissa bits (pinned on the GPU by
    tests/test_stream_gpu.py::test_tf32_operands_are_truncated), so with c = 1 - 2^-10
        | dot(trunc(z), trunc(e)) / c  -  dot(z, e) |  <=  kTf32Eps * |z| * |e|        (kTf32Eps = 1.03e-3 >= 2^-10 / c)
    for every z, e -- half of what the uncentred estimate dot(trunc(z), trunc(e)) can guarantee (2^-9).

Random, adversarial (all mantissas just below / exactly at a truncation step) and exactly representable operands.
"""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _consts():
    src = (ROOT / "ct_vae_b200" / "csrc" / "ctvq_common.cuh").read_text()
    eps = float(re.search(r"kTf32Eps\s*=\s*([0-9.eE+-]+)f", src).group(1))
    m = re.search(r"kTruncC\s*=\s*1\.0f\s*-\s*([0-9.eE+-]+)f", src)
    return eps, 1.0 - float(m.group(1))


def _trunc_tf32(x: np.ndarray) -> np.ndarray:
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _worst_ratio(z: np.ndarray, e: np.ndarray, c: float):
    """max over rows of |dot(trunc z, trunc e)/c - dot(z, e)| / (|z| |e|), and the same for the uncentred estimate (float64)."""
    zt, et = _trunc_tf32(z).astype(np.float64), _trunc_tf32(e).astype(np.float64)
    z64, e64 = z.astype(np.float64), e.astype(np.float64)
    exact = (z64 * e64).sum(-1)
    tc = (zt * et).sum(-1)
    norm = np.linalg.norm(z64, axis=-1) * np.linalg.norm(e64, axis=-1)
    return float((np.abs(tc / c - exact) / norm).max()), float((np.abs(tc - exact) / norm).max())


def test_constants_are_what_the_bound_needs():
    eps, c = _consts()
    assert c == 1.0 - 2.0 ** -10
    assert eps >= 2.0 ** -10 / c, "kTf32Eps must cover the half-width of (1-dz)(1-de) around c, divided by c"
    assert eps < 0.55 * 2.0 ** -9 * 1.05, "the centred window is meant to be about half the uncentred one"


@pytest.mark.parametrize("d", [32, 64, 128, 256])
def test_centred_bound_holds(d):
    eps, c = _consts()
    rng = np.random.default_rng(d)
    cases = []
    z = rng.standard_normal((4096, d)).astype(np.float32)
    e = (rng.standard_normal((4096, d)) * 0.5).astype(np.float32)
    cases.append((z, e))
    cases.append((z, (rng.uniform(-1 / 512, 1 / 512, (4096, d))).astype(np.float32)))  # init-scale codebook
    # adversarial: every mantissa just below the next tf32 step (maximal truncation loss, all products the same sign) ...
    hi = np.nextafter(np.float32(1.0 + 2.0 ** -10), np.float32(0.0))
    cases.append((np.full((4, d), hi, np.float32), np.full((4, d), hi, np.float32)))
    # ... and exactly representable operands (no loss at all: the centred estimate then OVER-shoots by 2^-10 / c)
    cases.append((np.full((4, d), 1.5, np.float32), np.full((4, d), -0.75, np.float32)))
    # mixed: half the channels lose the maximum, half lose nothing, signs chosen so the errors add up
    zm = np.where(np.arange(d) % 2 == 0, hi, np.float32(1.0)).astype(np.float32)[None].repeat(4, 0)
    cases.append((zm, zm.copy()))
    for z_, e_ in cases:
        centred, plain = _worst_ratio(z_, e_, c)
        assert centred <= eps, f"centred estimate off by {centred:.3e} |z||e| > kTf32Eps = {eps:.3e}"
        assert plain <= 2.0 ** -9
    # the adversarial case really needs the full uncentred 2^-9: centring is what buys the factor of two
    assert _worst_ratio(cases[2][0], cases[2][1], c)[1] > 1.9 * 2.0 ** -10
