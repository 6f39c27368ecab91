"""Host logic: the kernels' __host__ __device__ bodies (sigproc_device.cuh), replayed on the CPU with the
kernels' own views and table-construction order, against the oracle and the golden vectors.  CPU only.
What this cannot see (device rounding of the intrinsics, warp staging) is covered by the -m gpu tests."""
import numpy as np
import pytest

from conftest import assert_same, golden
from emu import Emu
import synth


@pytest.fixture(scope="module")
def emu(hostemu):
    return Emu(hostemu, 1)


def test_tables_match_golden(emu):
    for sps in (1, 4):
        emu.setup(sps)
        g = golden("tables_sps%d.npz" % sps)
        for t, k in enumerate(["cos", "sin", "rot", "revrot", "pulse"]):
            assert_same(emu.table(t), g[k], k)
        for i in range(8):
            assert_same(emu.table(5, i), g["mid_seq"][i], "mid_seq")
            assert_same(emu.table(6, i), g["mid_meta"][i], "mid_meta")
        assert_same(emu.table(7), g["rach_seq"], "rach_seq")
        assert_same(emu.table(8), g["rach_meta"], "rach_meta")
        assert_same(emu.table(9), g["lpf_rx"], "lpf_rx")
        assert_same(emu.table(10), g["lpf_tx"], "lpf_tx")
    emu.setup(1)


def test_closed_form_range_reduction_equals_loops(emu, oracle_port):
    """trig_lookup replaces the reference's while-loops by one exact subtraction / one rounded addition"""
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.uniform(-80, 80, 40000), rng.uniform(-7, 7, 20000), rng.uniform(-0.02, 0.02, 2000),
                         np.arange(-12, 13) * np.float32(np.pi), np.arange(-12, 13) * 2 * np.float32(np.pi),
                         [0.0, 0.01, -0.01, 1e-8, -1e-8, 6.2831855, -6.2831855]]).astype(np.float32)
    for x in xs:
        x = float(x)
        assert emu.lib.emu_sin_lookup(x) == oracle_port.sin_lookup(x), x
        assert emu.lib.emu_cos_lookup(x) == oracle_port.cos_lookup(x), x
        assert emu.lib.emu_sinc(x) == oracle_port.sinc(x), x


def test_sinc_grid_is_exact(emu):
    assert emu.lib.emu_check_sinc_grid() == 0


def test_grid_peak_detect_equals_general(emu, oracle_port):
    rng = np.random.default_rng(4)
    for n in (36, 156, 157, 26, 41):
        for _ in range(40):
            v = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
            k = rng.integers(0, n)
            v[k] += 6 * np.exp(1j * rng.uniform(0, 6.28))
            if k + 1 < n:
                v[k + 1] += rng.uniform(0, 6)
            ref = oracle_port.peak_detect(v)
            assert emu.peak_detect(v, grid=False) == ref
            assert emu.peak_detect(v, grid=True) == ref
    z = np.zeros(36, np.complex64)                      # all-zero input: maxIndex stays -1
    assert emu.peak_detect(z, grid=True) == oracle_port.peak_detect(z)
    e = np.zeros(36, np.complex64); e[0] = 3; e[35] = 3.5  # peaks on the edges
    assert emu.peak_detect(e, grid=True) == oracle_port.peak_detect(e)


def test_delay_vector(emu, oracle_port):
    rng = np.random.default_rng(5)
    v = (rng.standard_normal(156) + 1j * rng.standard_normal(156)).astype(np.complex64)
    for d in (0.0, 0.005, 0.3, 6.932, -2.75, -0.999, 3.0, -3.0, 1.5, 7 / 512, 5 / 512, 6 / 512, -1 + 300 / 512, 200.0, -200.0):
        assert_same(emu.delay_vector(v, d), oracle_port.delay_vector(v, d), "delay %r" % d)


def test_design_dfe_and_equalize(emu, oracle_port):
    rng = np.random.default_rng(6)
    for nchan, nf in ((6, 7), (4, 7), (2, 5), (1, 3), (7, 7)):
        ch = (rng.standard_normal(nchan) + 1j * rng.standard_normal(nchan)).astype(np.complex64) * 0.4
        ch[0] += 1
        ref = oracle_port.design_dfe(ch, 37.5, nf)
        got = emu.design_dfe(ch, 37.5, nf)
        assert_same(got[0], ref[0], "w"); assert_same(got[1], ref[1], "b")
        if (nchan, nf) == (6, 7):
            fx = emu.design_dfe(ch, 37.5, nf, fixed=True)
            assert_same(fx[0], ref[0], "w<7,5>"); assert_same(fx[1], ref[1], "b<7,5>")
            x = (rng.standard_normal(157) + 1j * rng.standard_normal(157)).astype(np.complex64)
            for toa in (0.0, 1.296875, -2.5, 3.0):
                a, b = emu.equalize(x, toa, *ref), oracle_port.equalize(x, toa, *ref)
                assert_same(a[0], b[0], "soft"); assert_same(a[1], b[1], "burst after")


@pytest.fixture(params=["ring", "rolling"])
def eq_tile(request, hostemu):
    """which equaliser kernel's tile policy the emulation replays: k_equalize_ring (shipped) or k_equalize_fast (the A/B);
    either way rows the kernel does not hold at that step read as poison"""
    hostemu.emu_set_eq_ring(1 if request.param == "ring" else 0)
    yield request.param
    hostemu.emu_set_eq_ring(1)


def test_demod_normal_kernel_logic_matches_golden(emu, eq_tile):
    g = golden("normal_sps1.npz")
    r = emu.rx_normal_batch(g["bursts"], g["lens"], g["tsc"])
    for k in ("flag", "amp", "toa", "chan", "off", "w", "b", "soft"):
        assert_same(r[k], g[k], k)


def test_demod_normal_kernel_logic_random(emu, oracle_port, eq_tile):
    mod = lambda b, gd: oracle_port.modulate(b, gd)  # noqa: E731
    bursts, lens, tsc, _ = synth.make_normal_batch(mod, 600, seed=21, noise_only=0.1, snr=(0, 30))
    a, b = emu.rx_normal_batch(bursts, lens, tsc), oracle_port.rx_normal_batch(bursts, lens, tsc, threads=2)
    for k in b:
        assert_same(a[k], b[k], k)
    # rule-based lengths (lens == NULL) and the energy gate
    a = emu.rx_normal_batch(bursts, None, tsc, gate_thr=600.0)
    gate = np.array([oracle_port.energy_detect(bursts[i, :lens[i]], 20, 600.0)[0] for i in range(len(lens))])
    assert (a["flag"] == (b["flag"] & gate)).all() and 0 < gate.sum() < gate.size


def test_demod_normal_kernel_logic_edge_cases(emu, oracle_port, eq_tile):
    bursts, lens, tsc = synth.make_edge_batch(oracle_port)
    a, b = emu.rx_normal_batch(bursts, lens, tsc), oracle_port.rx_normal_batch(bursts, lens, tsc)
    assert b["flag"].sum() > 10
    for k in b:
        assert_same(a[k], b[k], k)


def test_rach_kernel_logic(emu, oracle_port, eq_tile):
    g = golden("rach_sps1.npz")
    for tiles in (True, False):
        r = emu.rx_rach_batch(g["bursts"], g["lens"], tiles=tiles)
        for k in ("flag", "amp", "toa", "soft"):
            assert_same(r[k][..., :160], g[k], "%s tiles=%s" % (k, tiles))


def test_sps4_analyze_and_slicer(emu):
    emu.setup(4)
    g = golden("sps4.npz")
    n = g["rx"].shape[0]
    r = emu.analyze_batch(g["rx"], np.full(n, g["rx"].shape[1], np.int32), g["tsc"])
    assert_same(r["flag"], g["ok"].astype(np.int32), "flag"); assert_same(r["amp"], g["amp"], "amp")
    assert_same(r["toa"], g["toa"], "toa"); assert_same(r["chan"], g["chan"], "chan"); assert_same(r["off"], g["off"], "off")
    emu.setup(1)


def test_stream_kernels_logic(emu):
    g = golden("stream_sps1.npz")
    assert_same(emu.tx_resample_stream(g["stream_head"]), g["iq_head"], "TX resample")
    assert_same(emu.rx_resample_stream(g["raw_head"]), g["res_head"], "RX resample")
    res = g["res_head"]
    r = emu.rx_normal_batch(res, None, np.zeros(64, np.uint8), pitch=0)      # slot-stream addressing
    for k in ("flag", "amp", "toa", "soft"):
        assert_same(r[k], g[k], k)
    b = g["bits"][0]
    assert_same(emu.modulate(b, 9), g["stream_head"][:157], "modulate")


def test_tuned_rx_resampler_logic(emu, oracle_port):
    """period/phase formulation of the chunked reference loop: zero history, right-edge truncation of every chunk,
    tiles that are not a multiple of a chunk, a stream that does not fill the last tile"""
    g = golden("stream_sps1.npz")
    assert_same(emu.rx_resample_stream_v2(g["raw_head"]), g["res_head"], "v2 == golden")
    rng = np.random.default_rng(8)
    for nch in (1, 2, 3, 4, 7, 11, 36):
        raw = ((rng.standard_normal(nch * 864) + 1j * rng.standard_normal(nch * 864)) * 3000).astype(np.complex64)
        ref = oracle_port.rx_resample_stream(raw)
        assert_same(emu.rx_resample_stream_v2(raw), ref, "v2 %d chunks" % nch)
        if nch > 2:
            assert_same(emu.rx_resample_stream_v2(raw, has_history=True, offset=2 * 864), ref[2 * 585:], "v2 with history")


def test_tx_fused_chain(emu, oracle_best):
    """the fused TX kernel's arithmetic (modulate into the tile, 96-phase periods, quantise) == modulateBurst +
    pushBuffer of the oracle, with and without per-slot power scaling"""
    rng = np.random.default_rng(3)
    nb = 936
    bits = np.stack([synth.normal_burst_bits(rng, i % 8) for i in range(nb)])
    want = oracle_best.tx_resample_stream(oracle_best.modulate_stream(bits, threads=4), threads=4)
    got = emu.tx_fused(bits)
    assert np.array_equal(got, want.reshape(got.shape))
    ones = emu.tx_fused(bits, np.ones(nb, np.float32))
    assert np.array_equal(ones, got)
    half = emu.tx_fused(bits, np.full(nb, 0.1, np.float32))
    assert np.abs(half).max() < np.abs(got).max() * 0.11
