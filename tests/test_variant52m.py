"""The reference's second transceiver variant (Transceiver52M/sigProcLib.cpp): the two functions whose
arithmetic differs from the main variant -- windowed analyzeTrafficBurst (maxTOA) and the stride-4 energyDetect -- against
that variant compiled in place (oracle/_ref/libref52_oracle.so)."""
import os

import numpy as np
import pytest

import synth
from emu import Emu

HAVE52 = os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libref52_oracle.so"))
pytestmark = pytest.mark.skipif(not HAVE52, reason="Transceiver52M oracle not built (needs /root/reference)")


@pytest.fixture(scope="module")
def o52():
    from oracle.oracle import Oracle52
    return Oracle52(1)


def cases(oracle_best, n=160, seed=8):
    """normal bursts over the whole TOA window, some noise-only, all TSCs"""
    bursts, lens, tsc, _ = synth.make_normal_batch(oracle_best.modulate, n, seed=seed, max_delay=6.0, noise_only=0.15)
    rng = np.random.default_rng(seed)
    for i in range(0, n, 5):                          # early arrivals too (negative TOA)
        bursts[i] = np.roll(bursts[i], -int(rng.integers(1, 5)))
    return bursts, lens, tsc


def same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and np.array_equal(a, b), what


def run_analyze(fn_ref, fn_got, bursts, lens, tsc):
    nd = 0
    for max_toa in (0, 3, 5, 8, 20, 45):
        for i in range(bursts.shape[0]):
            x = bursts[i, :lens[i]]
            r = fn_ref(x, int(tsc[i]), 3.0, max_toa, True)
            g = fn_got(x, int(tsc[i]), 3.0, max_toa, True)
            where = "burst %d maxTOA %d" % (i, max_toa)
            assert r[0] == g[0], where
            same(g[1], r[1], where + " amp"); same(g[2], r[2], where + " toa")
            if r[0]:
                same(g[3], r[3], where + " chan"); same(g[4], r[4], where + " off")
                nd += 1
    return nd


def test_analyze_52m_hostemu(oracle_best, o52, hostemu):
    bursts, lens, tsc = cases(oracle_best)
    emu = Emu(hostemu)
    nd = run_analyze(o52.analyze, emu.analyze_52m, bursts, lens, tsc)
    assert nd > 300
    rng = np.random.default_rng(1)
    for _ in range(50):
        v = (rng.standard_normal(156) + 1j * rng.standard_normal(156)).astype(np.complex64) * rng.uniform(1, 500)
        thr = float(rng.uniform(1, 500))
        assert o52.energy_detect(v, 20, thr) == emu.energy_detect_52m(v, 20, thr)


@pytest.mark.gpu
def test_analyze_52m_gpu(oracle_best, o52, dsp):
    bursts, lens, tsc = cases(oracle_best, n=96)
    nd = run_analyze(o52.analyze, dsp.analyze_52m, bursts[:24], lens[:24], tsc[:24])
    assert nd > 40
    for max_toa in (3, 12):                           # batched entry point
        got = dsp.analyze_52m_host(bursts, lens, tsc, 3.0, max_toa, True)
        for i in range(bursts.shape[0]):
            r = o52.analyze(bursts[i, :lens[i]], int(tsc[i]), 3.0, max_toa, True)
            assert bool(got["flag"][i]) == r[0]
            same(got["amp"][i], r[1], "amp"); same(got["toa"][i], r[2], "toa")
            if r[0]:
                same(got["chan"][i], r[3], "chan"); same(got["off"][i], r[4], "off")
    rng = np.random.default_rng(1)
    for _ in range(10):
        v = (rng.standard_normal(156) + 1j * rng.standard_normal(156)).astype(np.complex64) * rng.uniform(1, 500)
        thr = float(rng.uniform(1, 500))
        assert o52.energy_detect(v, 20, thr) == dsp.energy_detect_52m(v, 20, thr)
