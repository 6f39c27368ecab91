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
    # the tile-staged kernel's layout: only the search window is in the tile, the rest is poison
    assert run_analyze(o52.analyze, emu.analyze_52m_tiled, bursts, lens, tsc) == nd
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


# ---- that variant's caller policy (Transceiver52M/Transceiver.cpp:268-404) and transmit scaling (:74, :111) ----
CHAN_TYPES = [[1, 1, 5, 4, 0, 7, 2, 8], [5, 1, 1, 1, 6, 3, 1, 1], [1, 1, 1, 1, 1, 1, 1, 1]]
TSC = [2, 5, 0]
NFRAMES, FN0 = 80, 2715600                        # crosses the hyperframe wrap


@pytest.fixture(scope="module")
def batch52(oracle_best):
    return synth.make_trx_batch(oracle_best.modulate, oracle_best.expected_corr_type, NFRAMES, TSC, CHAN_TYPES, fn0=FN0, seed=52)


def oracle52_pull(o52, bursts, max_delay, split):
    A = len(TSC)
    b4 = bursts.reshape(NFRAMES, A, 8, -1)
    valid = np.zeros((NFRAMES, A, 8), np.int32)
    dg = np.zeros((NFRAMES, A, 8, 158), np.uint8)
    states = []
    for a in range(A):
        st = o52.trx_new(TSC[a], CHAN_TYPES[a], FN0 - 3)
        for lo, hi in split:
            v, d = o52.trx_pull(st, np.ascontiguousarray(b4[lo:hi, a]).reshape((hi - lo) * 8, -1), FN0 + lo, max_delay)
            valid[lo:hi, a] = v.reshape(hi - lo, 8)
            dg[lo:hi, a] = d[:, :158].reshape(hi - lo, 8, 158)
        states.append(st)
    return valid.reshape(-1), dg.reshape(-1, 158), np.concatenate(states)


def check_policy(pull, new_state, get_state, bursts, o52, where):
    from test_trx_policy import check_state
    A = len(TSC)
    for max_delay in (0, 1, 2, 4, 9):             # 0, 1: no equaliser (needDFE false); >= 2: DFE with a +-max_delay search
        for split in ([(0, NFRAMES)], [(0, 29), (29, 30), (30, NFRAMES)]):
            v1, d1, s1 = oracle52_pull(o52, bursts, max_delay, split)
            st = new_state(max_delay)
            v2, d2 = np.zeros_like(v1), np.zeros_like(d1)
            for lo, hi in split:
                v, d = pull(st, bursts[lo * A * 8:hi * A * 8], FN0 + lo, max_delay)
                v2[lo * A * 8:hi * A * 8] = v
                d2[lo * A * 8:hi * A * 8] = d[:, :158]
            assert np.array_equal(v1, v2), (where, max_delay)
            bad = np.nonzero((d1 != d2).any(axis=1))[0]
            assert bad.size == 0, (where, max_delay, bad[:10], d1[bad[:1]], d2[bad[:1]])
            check_state(get_state(st), s1, "%s maxdly %d" % (where, max_delay))
            assert 0.15 < v1.mean() < 0.9
            if max_delay <= 1:
                assert not s1["have"].any()       # that mode never caches a channel estimate


def test_policy_52m_hostemu(oracle_best, o52, hostemu, batch52):
    emu = Emu(hostemu)
    check_policy(emu.trx_pull_52m, lambda md: emu.trx_new(TSC, CHAN_TYPES, FN0 - 3), lambda st: st, batch52, o52, "hostemu")


@pytest.mark.gpu
def test_policy_52m_gpu(oracle_best, o52, dsp, batch52):
    made = []

    def new_state(md):
        trx = dsp.trx_create(TSC, CHAN_TYPES, FN0 - 3)
        dsp.trx_set_variant_52m(trx, True, md)
        made.append(trx)
        return trx
    check_policy(lambda trx, b, fn, md: dsp.trx_pull_host(trx, b, fn), new_state, dsp.trx_state, batch52, o52, "gpu")
    # switching a trx back restores the main variant's policy
    trx = made[-1]
    dsp.trx_set_variant_52m(trx, False, 0)
    so = oracle_best.trx_new(TSC[0], CHAN_TYPES[0], FN0 - 3)
    t2 = dsp.trx_create(TSC, CHAN_TYPES, FN0 - 3)
    v_main, d_main = dsp.trx_pull_host(t2, batch52, FN0)
    t3 = dsp.trx_create(TSC, CHAN_TYPES, FN0 - 3)
    dsp.trx_set_variant_52m(t3, True, 5)
    dsp.trx_set_variant_52m(t3, False, 0)
    v_back, d_back = dsp.trx_pull_host(t3, batch52, FN0)
    assert np.array_equal(v_main, v_back) and np.array_equal(d_main, d_back)
    for t in made + [t2, t3]:
        dsp.trx_destroy(t)


@pytest.mark.gpu
def test_tx_datagrams_52m_gpu(o52, dsp):
    """modulate-time scaling 13500 * pow(10, -RSSI/10) (fillers: 13500) and the symbol-rate radio's short casts"""
    rng = np.random.default_rng(7)
    nframes, fn0 = 23, 2715640
    n = 150
    dg = np.zeros((n, 154), np.uint8)
    dg[:, 0] = rng.integers(0, 8, n)
    dg[:5, 0] = [8, 255, 3, 3, 200]                                  # bad timeslots are dropped
    fns = (fn0 + rng.integers(-3, nframes + 3, n)) % (2048 * 26 * 51)
    for k in range(4):
        dg[:, 1 + k] = (fns >> ((3 - k) * 8)) & 0xff
    dg[:, 5] = rng.integers(0, 40, n)
    dg[7, 5] = 250                                                   # a negative RSSI byte (char): pow(10, +)
    dg[:, 6:154] = rng.integers(0, 2, (n, 148))
    filler = rng.integers(0, 2, 148).astype(np.uint8)
    for fl in (filler, None):
        want, wp = o52.tx_datagrams(dg, fn0, nframes, fl)
        got, gp = dsp.tx_datagrams_52m_host(dg, fn0, nframes, fl)
        assert wp == gp and 0 < gp < n
        same(got, want, "52M tx datagrams, filler %s" % (fl is not None))


def test_policy_52m_hostemu_matches_golden(oracle_best, hostemu):
    """the second variant's policy against the fixture the compiled 52M reference wrote (oracle/gen_golden_r3.py)"""
    import hashlib
    from conftest import golden
    from oracle.oracle import Oracle
    from test_trx_policy import check_state
    g = golden("trx52_sps1.npz")
    bursts = synth.make_trx_batch(oracle_best.modulate, oracle_best.expected_corr_type, NFRAMES, TSC, CHAN_TYPES, fn0=FN0, seed=52)
    if hashlib.sha1(bursts.tobytes()).digest() != g["sha1"].tobytes():
        pytest.skip("regenerated inputs differ from the fixture's (different modulator build)")
    emu = Emu(hostemu)
    A = len(TSC)
    for md in (1, 4):
        st = emu.trx_new(TSC, CHAN_TYPES, FN0 - 3)
        v2, d2 = np.zeros_like(g["valid%d" % md]), np.zeros_like(g["dgram%d" % md])
        for lo, hi in [(0, 29), (29, 30), (30, NFRAMES)]:
            v, d = emu.trx_pull_52m(st, bursts[lo * A * 8:hi * A * 8], FN0 + lo, md)
            v2[lo * A * 8:hi * A * 8] = v
            d2[lo * A * 8:hi * A * 8] = d[:, :158]
        assert np.array_equal(v2, g["valid%d" % md]) and np.array_equal(d2, g["dgram%d" % md])
        check_state(st, g["state%d" % md].reshape(-1).view(Oracle.TRX_STATE_DTYPE), "52M golden maxdly %d" % md)
