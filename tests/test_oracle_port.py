"""The oracle itself: the plain-C port against the golden vectors (always) and against the compiled
reference call by call (when oracle/_ref is present).  CPU only."""
import numpy as np
import pytest

from conftest import assert_same, golden
from oracle import oracle as orc
import synth

needs_ref = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref/libref_oracle.so not built here")


def _tables(o):
    d = {}
    for t, name in enumerate(["cos", "sin", "rot", "revrot", "pulse"]):
        d[name] = o.table(t)
    d["mid_seq"] = np.stack([o.table(5, i) for i in range(8)])
    d["mid_meta"] = np.stack([o.table(6, i) for i in range(8)])
    d["rach_seq"] = o.table(7)
    d["rach_meta"] = o.table(8)
    d["lpf_rx"] = o.table(9)
    d["lpf_tx"] = o.table(10)
    return d


@pytest.mark.parametrize("sps", [1, 4])
def test_port_tables_match_golden(sps):
    o = orc.Oracle("port", sps=sps)
    g = golden("tables_sps%d.npz" % sps)
    for k, v in _tables(o).items():
        assert_same(v, g[k], "table %s sps %d" % (k, sps))
    # SURVEY 8(c) survey-time probe values
    if sps == 1:
        assert np.allclose(o.table(4).real, [0.182762086, 0.966020703, 0.182762086], rtol=0, atol=1e-9)
        assert [float(m[0]) for m in g["mid_meta"]] == [8.001953125, 7.998046875, 8.001953125, 7.998046875,
                                                        7.998046875, 7.998046875, 8.001953125, 8.001953125]
        assert float(g["rach_meta"][0]) == 20.0
    o.setup(1)


def test_port_config1_matches_golden(oracle_port):
    o, g = oracle_port, golden("config1_sps1.npz")
    tx = o.modulate(g["bits"], 8)
    assert_same(tx, g["tx"], "modulateBurst")
    assert_same(o.delay_vector(tx, 6.932), g["delayed"], "delayVector")
    assert_same(o.convolve(g["delayed"], np.array([9000, 3600, 0, 0], np.complex64), orc.NO_DELAY), g["chan_out"], "convolve")
    ok, amp, toa, chan, off = o.analyze(g["rx"], 0, 8.0, request=True)
    assert ok == bool(g["ok"]) and amp == g["amp"] and toa == float(g["toa"]) and off == float(g["off"])
    assert_same(chan, g["chan"], "channel")
    assert_same(o.demodulate(g["rx"], amp, toa), g["soft_slicer"], "demodulateBurst")
    w, b = o.design_dfe(chan, 1.0 / 0.001, 7)
    assert_same(w, g["w"], "w"); assert_same(b, g["b"], "b")
    soft, after = o.equalize(g["rx"], toa - off, w, b)
    assert_same(soft, g["soft_dfe"], "equalizeBurst"); assert_same(after, g["rx_after"], "burst after equalize")
    pk, idx, avg = o.peak_detect(o.correlate(g["rx"][56:92], o.table(5, 0), orc.NO_DELAY))
    assert pk == g["pk"] and idx == float(g["pidx"]) and avg == float(g["pavg"])
    assert_same(o.convolve(g["rx"], w, orc.FULL_SPAN), g["full"], "FULL_SPAN convolve")
    assert ((soft[:148] > 0.5) == g["bits"].astype(bool)).all()


def test_port_normal_batch_matches_golden(oracle_port):
    g = golden("normal_sps1.npz")
    r = oracle_port.rx_normal_batch(g["bursts"], g["lens"], g["tsc"], threads=2)
    for k in ("flag", "amp", "toa", "chan", "off", "w", "b", "soft"):
        assert_same(r[k], g[k], k)
    assert 0 < g["flag"].sum() < g["flag"].size          # the fixture holds detections and rejections


def test_port_rach_batch_matches_golden(oracle_port):
    g = golden("rach_sps1.npz")
    r = oracle_port.rx_rach_batch(g["bursts"], g["lens"], threads=2)
    for k in ("flag", "amp", "toa", "soft"):
        assert_same(r[k], g[k], k)


def test_port_stream_matches_golden(oracle_port):
    o, g = oracle_port, golden("stream_sps1.npz")
    assert_same(o.modulate_stream(g["bits"])[:20 * 585], g["stream_head"][:o.modulate_stream(g["bits"]).size][:20 * 585], "modulate stream")
    assert_same(o.tx_resample_stream(g["stream_head"]), g["iq_head"], "TX resample + int16")
    res = o.rx_resample_stream(g["raw_head"])
    assert_same(res, g["res_head"], "RX resample")
    nb = 64
    d = o.rx_stream_demod(res, nb, np.zeros(nb, np.uint8))
    for k in ("flag", "amp", "toa", "soft"):
        assert_same(d[k], g[k], k)
    assert float(g["ber"]) == 0.0


def test_port_sps4_matches_golden():
    o = orc.Oracle("port", sps=4)
    g = golden("sps4.npz")
    for i in range(g["rx"].shape[0]):
        ok, amp, toa, chan, off = o.analyze(g["rx"][i], int(g["tsc"][i]), 3.0, request=True)
        assert ok == bool(g["ok"][i]) and amp == g["amp"][i] and toa == g["toa"][i] and off == g["off"][i]
        assert_same(chan, g["chan"][i], "chan")
        assert_same(o.demodulate(g["rx"][i], amp, toa), g["soft"][i], "soft")
    o.setup(1)


@needs_ref
def test_port_equals_reference_randomised():
    """call-by-call against the compiled reference on fresh random inputs (not only the fixtures)"""
    R, P = orc.Oracle("ref"), orc.Oracle("port")
    rng = np.random.default_rng(7)
    for n, lb in ((36, 16), (156, 41), (157, 7), (20, 21), (5, 9)):
        a = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        b = (rng.standard_normal(lb) + 1j * rng.standard_normal(lb)).astype(np.complex64)
        for span in range(5):
            for ar in (False, True):
                for br in (False, True):
                    assert_same(R.convolve(a, b, span, ar, br), P.convolve(a, b, span, ar, br), "convolve")
                    assert_same(R.correlate(a, b, span, ar, br), P.correlate(a, b, span, ar, br), "correlate")
        for d in (0.0, 0.005, 0.3, 6.932, -2.75, -0.999, 200.0, -200.0):
            assert_same(R.delay_vector(a, d), P.delay_vector(a, d), "delay %r" % d)
        assert R.peak_detect(a) == P.peak_detect(a)
        for ix in (-3.2, 0.0, 4.37, n - 1.5, n + 4.0):
            assert R.interpolate_point(a, ix) == P.interpolate_point(a, ix)
        assert R.energy_detect(a, 20, 0.9) == P.energy_detect(a, 20, 0.9)
    xs = np.concatenate([rng.uniform(-40, 40, 20000), rng.uniform(-0.02, 0.02, 2000), [0.0, 0.01, -0.01, 6.2831855]])
    for x in xs.astype(np.float32):
        assert R.sinc(float(x)) == P.sinc(float(x)) or (np.isnan(R.sinc(float(x))) and np.isnan(P.sinc(float(x))))
    for nchan, nf in ((6, 7), (4, 7), (2, 5), (1, 3), (7, 7)):
        ch = (rng.standard_normal(nchan) + 1j * rng.standard_normal(nchan)).astype(np.complex64) * 0.4
        ch[0] += 1
        for a, b in zip(R.design_dfe(ch, 37.5, nf), P.design_dfe(ch, 37.5, nf)):
            assert_same(a, b, "designDFE")
    x = (rng.standard_normal(1056) + 1j * rng.standard_normal(1056)).astype(np.complex64)
    assert_same(R.resample(x, 65, 96, 0), P.resample(x, 65, 96, 0), "resample rx")
    assert_same(R.resample(x[:715], 96, 65, 1), P.resample(x[:715], 96, 65, 1), "resample tx")
    mod = lambda b, g: R.modulate(b, g)  # noqa: E731
    bursts, lens, tsc, _ = synth.make_normal_batch(mod, 400, seed=11, noise_only=0.1)
    a, b = R.rx_normal_batch(bursts, lens, tsc, threads=2), P.rx_normal_batch(bursts, lens, tsc, threads=2)
    for k in a:
        assert_same(a[k], b[k], "rx_normal " + k)
    rb, rl, _, _ = synth.make_rach_batch(mod, 200, seed=12)
    a, b = R.rx_rach_batch(rb, rl, threads=2), P.rx_rach_batch(rb, rl, threads=2)
    for k in a:
        assert_same(a[k], b[k], "rx_rach " + k)


@needs_ref
def test_reference_oob_tap_is_negligible():
    """createLPF(.,961,.) reads sendLPF_961[960], one past the table; we define it as 0"""
    R = orc.Oracle("ref")
    assert abs(float(R.table(orc.T_OOB)[0])) < 1e-30
    assert abs(float(R.table(orc.T_LPF_RX)[960])) < 1e-30


@pytest.mark.parametrize("kind", ["port", pytest.param("ref", marks=needs_ref)])
def test_wire_format_glue_of_the_oracles(kind):
    """the int16 ingest (unUSRPifyVector + pullBuffer resample) equals the float resample of the widened samples, in both
    I/Q orders and in any threading, and the soft-byte conversion is (char) round(soft*255.0) (Transceiver.cpp:667-669)"""
    o = orc.Oracle(kind, sps=1)
    rng = np.random.default_rng(21)
    iq = rng.integers(-32768, 32768, (7 * 864, 2)).astype(np.int16)
    for flip in (False, True):
        wide = (iq[:, 1] + 1j * iq[:, 0] if flip else iq[:, 0] + 1j * iq[:, 1]).astype(np.complex64)
        want = o.rx_resample_stream(wide)
        assert_same(o.rx_resample_stream_i16(iq, flip), want, "i16 ingest, flip=%s" % flip)
        assert_same(o.rx_resample_stream_i16(iq, flip, threads=3), want, "i16 ingest threaded")
    soft = rng.random((33, 160)).astype(np.float32)
    soft[0, :4] = [0.0, 1.0, 0.5, np.float32(0.5) - np.float32(2 ** -25)]
    got = o.soft_to_wire(soft, threads=2)
    want = np.floor(soft[:, :148].astype(np.float64) * 255.0 + 0.5).astype(np.uint8)     # round half away from zero, x >= 0
    assert_same(got, want, "soft bytes")


@needs_ref
def test_port_wire_glue_matches_reference():
    p, r = orc.Oracle("port", sps=1), orc.Oracle("ref", sps=1)
    rng = np.random.default_rng(22)
    iq = rng.integers(-20000, 20000, (5 * 864, 2)).astype(np.int16)
    assert_same(p.rx_resample_stream_i16(iq, True), r.rx_resample_stream_i16(iq, True), "i16 ingest port vs ref")
    soft = rng.random((64, 148)).astype(np.float32)
    assert_same(p.soft_to_wire(soft), r.soft_to_wire(soft), "soft bytes port vs ref")
