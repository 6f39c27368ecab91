"""Parity of the CUDA path with the reference, through the C ABI (libbtsdsp.so) on cuda:0.

Checker: oracle_best = the compiled reference (oracle/_ref, travels with the repo) when present, else the
plain-C port.  Bar (BASELINE.json north_star): detection flags, integer TOA and hard bits exact; soft bits,
amplitudes and fractional TOA within 1e-4 relative.  The kernels are designed to be bit-identical, so the
tests assert exact equality of every float and would only fall back to the stated tolerance knowingly
(REL_TOL below is the contract's alarm threshold, not used unless EXACT is switched off).
"""
import numpy as np
import pytest

from conftest import assert_same, golden
import synth

pytestmark = pytest.mark.gpu

EXACT = True
REL_TOL = 1e-4          # north_star tolerance for soft bits / amplitudes / fractional TOA


def same(a, b, what):
    if EXACT:
        assert_same(a, b, what)
    else:
        a, b = np.asarray(a), np.asarray(b)
        scale = max(float(np.abs(b).max()), 1e-30)
        assert np.all(np.abs(a - b) <= REL_TOL * np.maximum(np.abs(b), 1e-6 * scale)), what


def check_batch(got, ref, keys):
    same(got["flag"], ref["flag"], "detection flags")
    same(np.rint(got["toa"]), np.rint(ref["toa"]), "integer TOA")
    if "soft" in keys:
        same(got["soft"][:, :148] > 0.5, ref["soft"][:, :148] > 0.5, "hard bits")
    for k in keys:
        same(got[k], ref[k][..., :got[k].shape[-1]] if got[k].ndim > 1 else ref[k], k)


# ---------------------------------------------------------------------------------------------- tables
@pytest.mark.parametrize("sps", [1, 4])
def test_tables(dsp, dsp4, sps):
    d = dsp if sps == 1 else dsp4
    g = golden("tables_sps%d.npz" % sps)
    for t, k in enumerate(["cos", "sin", "rot", "revrot", "pulse"]):
        assert_same(d.table(t), g[k], k)
    for i in range(8):
        assert_same(d.table(5, i), g["mid_seq"][i], "mid_seq %d" % i)
        assert_same(d.table(6, i), g["mid_meta"][i], "mid_meta %d" % i)
    assert_same(d.table(7), g["rach_seq"], "rach_seq")
    assert_same(d.table(8), g["rach_meta"], "rach_meta")
    assert_same(d.table(9), g["lpf_rx"], "lpf_rx")
    assert_same(d.table(10), g["lpf_tx"], "lpf_tx")


# ------------------------------------------------------------------------- layer 1: sigProcLib.h surface
def test_config1_flow_matches_golden(dsp):
    """Transceiver/sigProcLibTest.cpp's flow (SURVEY config 1) at sps 1, every intermediate pinned"""
    g = golden("config1_sps1.npz")
    from openbts_ttsou_b200 import NO_DELAY, FULL_SPAN
    tx = dsp.modulate(g["bits"], 8)
    same(tx, g["tx"], "modulateBurst")
    same(dsp.delay_vector(tx, 6.932), g["delayed"], "delayVector")
    same(dsp.convolve(g["delayed"], np.array([9000, 3600, 0, 0], np.complex64), NO_DELAY), g["chan_out"], "convolve")
    ok, amp, toa, chan, off = dsp.analyze(g["rx"], 0, 8.0, request=True)
    assert ok == bool(g["ok"]) and amp == g["amp"] and toa == float(g["toa"]) and off == float(g["off"])
    same(chan, g["chan"], "channel")
    same(dsp.demodulate(g["rx"], amp, toa), g["soft_slicer"], "demodulateBurst")
    w, b = dsp.design_dfe(chan, 1.0 / 0.001, 7)
    same(w, g["w"], "w"); same(b, g["b"], "b")
    soft, after = dsp.equalize(g["rx"], toa - off, w, b)
    same(soft, g["soft_dfe"], "equalizeBurst"); same(after, g["rx_after"], "burst after equalizeBurst")
    pk, idx, avg = dsp.peak_detect(dsp.correlate(g["rx"][56:92], dsp.table(5, 0), NO_DELAY))
    assert pk == g["pk"] and idx == float(g["pidx"]) and avg == float(g["pavg"])
    same(dsp.convolve(g["rx"], w, FULL_SPAN), g["full"], "FULL_SPAN convolve")
    assert ((soft[:148] > 0.5) == g["bits"].astype(bool)).all()


def test_single_vector_functions_randomised(dsp, oracle_best):
    o = oracle_best
    rng = np.random.default_rng(17)
    for n, lb in ((36, 16), (156, 41), (157, 7), (20, 21), (5, 9), (1000, 33)):
        a = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        b = (rng.standard_normal(lb) + 1j * rng.standard_normal(lb)).astype(np.complex64)
        for span in range(5):
            for ar, br in ((False, False), (True, False), (False, True), (True, True)):
                same(dsp.convolve(a, b, span, ar, br), o.convolve(a, b, span, ar, br), "convolve")
                same(dsp.correlate(a, b, span, ar, br), o.correlate(a, b, span, ar, br), "correlate")
        for d in (0.0, 0.005, 0.3, 6.932, -2.75, -0.999, 3.0, 7 / 512, 5 / 512, 200.0, -2000.0):
            same(dsp.delay_vector(a, d), o.delay_vector(a, d), "delay %r" % d)
        assert dsp.peak_detect(a) == o.peak_detect(a)
        for ix in (-3.2, 0.0, 4.37, n - 1.5, n + 4.0):
            assert dsp.interpolate_point(a, ix) == o.interpolate_point(a, ix)
        assert dsp.energy_detect(a, 20, 0.9) == o.energy_detect(a, 20, 0.9)
        same(dsp.scale_vector(a, 0.3 - 1.7j), o.scale_vector(a, 0.3 - 1.7j), "scaleVector")
        same(dsp.scale_vector(a, 0.3 - 1.7j, True), o.scale_vector(a, 0.3 - 1.7j, True), "scaleVector real")
    for nchan, nf in ((6, 7), (4, 7), (2, 5), (1, 3), (7, 7)):
        ch = (rng.standard_normal(nchan) + 1j * rng.standard_normal(nchan)).astype(np.complex64) * 0.4
        ch[0] += 1
        for x, y in zip(dsp.design_dfe(ch, 37.5, nf), o.design_dfe(ch, 37.5, nf)):
            same(x, y, "designDFE %d/%d" % (nchan, nf))
    x = (rng.standard_normal(1056) + 1j * rng.standard_normal(1056)).astype(np.complex64)
    same(dsp.resample(x, 65, 96, 0), o.resample(x, 65, 96, 0), "polyphaseResampleVector rx")
    same(dsp.resample(x[:715], 96, 65, 1), o.resample(x[:715], 96, 65, 1), "polyphaseResampleVector tx")
    for guard in (0, 8, 9):
        bits = rng.integers(0, 2, 148).astype(np.uint8)
        same(dsp.modulate(bits, guard), o.modulate(bits, guard), "modulateBurst")


def test_error_behaviour(dsp, dsp4):
    import openbts_ttsou_b200 as pkg
    x = np.ones(156, np.complex64)
    with pytest.raises(pkg.BtsDspError):           # DFE path is undefined off symbol rate
        dsp4.equalize(np.ones(624, np.complex64), 0.0, np.ones(7, np.complex64), np.ones(5, np.complex64))
    with pytest.raises(pkg.BtsDspError):
        dsp.analyze(x[:50], 0, 3.0)                # shorter than the correlation window
    with pytest.raises(pkg.BtsDspError):
        dsp.analyze(x, 9, 3.0)                     # TSC out of range (the reference asserts)
    ok, amp, toa = dsp.detect_rach(np.zeros(156, np.complex64), 5.0)
    assert not ok and amp == 0


# -------------------------------------------------------------------------------- batched, host buffers
def test_normal_batch_matches_golden(dsp):
    g = golden("normal_sps1.npz")
    r = dsp.demod_normal_host(g["bursts"], g["lens"], g["tsc"])
    check_batch(r, g, ("amp", "toa", "chan", "off", "w", "b", "soft"))


def test_normal_batch_config4_random(dsp, oracle_best):
    """1024 ARFCN x 8 TS, mixed TSC 0-7 (hits SURVEY F6), multipath, SNR 10-30 dB, some empty slots"""
    mod = lambda b, gd: oracle_best.modulate(b, gd)  # noqa: E731
    bursts, lens, tsc, bits = synth.make_normal_batch(mod, 8192, seed=0xC4, noise_only=0.05)
    ref = oracle_best.rx_normal_batch(bursts, lens, tsc, threads=8)
    r = dsp.demod_normal_host(bursts, lens, tsc)
    check_batch(r, ref, ("amp", "toa", "chan", "off", "w", "b", "soft"))
    r2 = dsp.demod_normal_host(bursts, None, tsc, debug=False, soft_pitch=148)    # rule-based lengths, packed soft
    same(r2["soft"], ref["soft"][:, :148], "soft (pitch 148)")
    det = ref["flag"] == 1
    good = det & np.isin(tsc, (0, 2, 6, 7))
    ber = ((r["soft"][good, :148] > 0.5) != bits[good].astype(bool)).mean()
    assert det.mean() > 0.9 and ber < 0.01


def test_normal_batch_edge_cases(dsp, oracle_best):
    bursts, lens, tsc = synth.make_edge_batch(oracle_best)
    n = len(lens)
    ref = oracle_best.rx_normal_batch(bursts, lens, tsc)
    assert ref["flag"].sum() > 10
    r = dsp.demod_normal_host(bursts, lens, tsc)
    check_batch(r, ref, ("amp", "toa", "chan", "off", "w", "b", "soft"))
    one = dsp.demod_normal_host(bursts[5:6], lens[5:6], tsc[5:6])                  # batch of one
    same(one["soft"], ref["soft"][5:6], "batch of 1")
    odd = dsp.demod_normal_host(bursts, lens, tsc, debug=False, soft_pitch=157)    # unaligned soft rows
    same(odd["soft"], ref["soft"][:, :157], "soft pitch 157")
    gate = dsp.demod_normal_host(bursts, lens, tsc, gate_thr=500.0)
    eg = np.array([oracle_best.energy_detect(bursts[i, :lens[i]], 20, 500.0)[0] for i in range(n)])
    same(gate["flag"], ref["flag"] & eg, "energy gate")
    same(gate["soft"], ref["soft"] * eg[:, None], "energy gate soft")


def test_rach_batch(dsp, dsp4, oracle_best):
    # dsp4 (an sps = 4 context on the same device) exists by now: per-device __constant__ tables must not be clobbered
    assert dsp4.sps == 4
    g = golden("rach_sps1.npz")
    r = dsp.rach_host(g["bursts"], g["lens"])
    check_batch(r, g, ("amp", "toa", "soft"))
    mod = lambda b, gd: oracle_best.modulate(b, gd)  # noqa: E731
    bursts, lens, bits, delays = synth.make_rach_batch(mod, 4096, seed=0xC3)       # config 3: TOA 0-63, SNR -5..20
    ref = oracle_best.rx_rach_batch(bursts, lens, threads=8)
    r = dsp.rach_host(bursts, lens)
    check_batch(r, ref, ("amp", "toa", "soft"))
    only = dsp.rach_host(bursts, lens, demod=False)
    same(only["flag"], ref["flag"], "detect only")
    assert 0.3 < ref["flag"].mean() < 1.0
    det = ref["flag"] == 1
    err = np.abs(ref["toa"][det] - delays[det])               # low-SNR false alarms land anywhere; most do not
    assert np.median(err) < 0.25 and (err < 1.5).mean() > 0.9


def test_sps4_functions(dsp4):
    g = golden("sps4.npz")
    for i in range(g["rx"].shape[0]):
        ok, amp, toa, chan, off = dsp4.analyze(g["rx"][i], int(g["tsc"][i]), 3.0, request=True)
        assert ok == bool(g["ok"][i]) and amp == g["amp"][i] and toa == g["toa"][i] and off == g["off"][i]
        same(chan, g["chan"][i], "chan")
        same(dsp4.demodulate(g["rx"][i], amp, toa), g["soft"][i], "demodulateBurst sps4")
    same(dsp4.modulate(golden("config1_sps1.npz")["bits"], 8).size, 4 * 156, "modulated length at sps 4")


# ------------------------------------------------------------------------------------- streams (config 2/5)
def test_stream_kernels_match_golden(dsp):
    import torch
    g = golden("stream_sps1.npz")
    dev = torch.device("cuda:0")
    x = torch.from_numpy(g["stream_head"].view(np.float32).copy()).to(dev)
    iq = torch.zeros(20 * 864 * 2, dtype=torch.int16, device=dev)
    dsp.resample_tx_dev(x, 20, iq)
    same(iq.cpu().numpy().reshape(-1, 2), g["iq_head"], "TX resample")
    raw = torch.from_numpy(g["raw_head"].view(np.float32).copy()).to(dev)
    nch = g["raw_head"].size // 864
    res = torch.zeros(nch * 585 * 2, dtype=torch.float32, device=dev)
    dsp.resample_rx_dev(raw, nch, res)
    same(res.cpu().numpy().view(np.complex64), g["res_head"], "RX resample")
    # a later part of the stream with real history == the same samples computed from the start
    res2 = torch.zeros((nch - 3) * 585 * 2, dtype=torch.float32, device=dev)
    dsp.resample_rx_dev(raw[3 * 864 * 2:], nch - 3, res2, has_history=True)
    same(res2.cpu().numpy().view(np.complex64), g["res_head"][3 * 585:], "RX resample with history")
    nb = 64
    tsc = torch.zeros(nb, dtype=torch.uint8, device=dev)
    flag = torch.zeros(nb, dtype=torch.int32, device=dev)
    amp = torch.zeros(nb * 2, dtype=torch.float32, device=dev)
    toa = torch.zeros(nb, dtype=torch.float32, device=dev)
    soft = torch.zeros(nb * 160, dtype=torch.float32, device=dev)
    dsp.demod_normal_dev(res, 0, tsc, nb, flag, amp, toa, soft, 160)       # pitch 0 = slot-stream addressing
    torch.cuda.synchronize()
    same(flag.cpu().numpy(), g["flag"], "flag"); same(amp.cpu().numpy().view(np.complex64), g["amp"], "amp")
    same(toa.cpu().numpy(), g["toa"], "toa"); same(soft.cpu().numpy().reshape(nb, 160), g["soft"], "soft")
    bits = torch.from_numpy(g["bits"].copy()).to(dev)
    st = torch.zeros(16 * 625 * 2, dtype=torch.float32, device=dev)
    dsp.modulate_dev(bits, 148, 64, st, 0)
    same(st.cpu().numpy().view(np.complex64), g["stream_head"][:16 * 625], "modulate to slot stream")


def test_stream_end_to_end_host_roundtrip(dsp, oracle_best):
    """TX (bits -> modulate -> 96/65 resample -> int16) then RX (65/96 resample -> slot cut -> demod), through the
    host-buffer entry points, checked (a) bit-exactly against the oracle on every burst and (b) by the round-trip
    property: clean TSC-0 bursts come back with zero bit errors."""
    rng = np.random.default_rng(0xC2)
    nb = 936 * 4                                    # 4 blocks of 117 frames = 1000 chunks
    bits = np.stack([synth.normal_burst_bits(rng, 0) for _ in range(nb)])
    iq = np.zeros((1000 * 864, 2), np.int16)
    dsp.tx_stream_host(bits, nb, iq)
    same(iq, oracle_best.tx_resample_stream(oracle_best.modulate_stream(bits, threads=4), threads=4), "TX stream")
    raw = (iq[:, 0] + 1j * iq[:, 1]).astype(np.complex64)
    raw = (raw + 30.0 * (rng.standard_normal(raw.size) + 1j * rng.standard_normal(raw.size))).astype(np.complex64)
    tsc = np.zeros(nb, np.uint8)
    flag, amp, toa = np.zeros(nb, np.int32), np.zeros(nb, np.complex64), np.zeros(nb, np.float32)
    soft = np.zeros((nb, 148), np.float32)
    dsp.rx_stream_host(raw, 1000, tsc, nb, flag, amp, toa, soft, 148)
    ref = oracle_best.rx_stream_demod(oracle_best.rx_resample_stream(raw, threads=4), nb, tsc, threads=8)
    same(flag, ref["flag"], "flag"); same(amp, ref["amp"], "amp"); same(toa, ref["toa"], "toa")
    same(soft, ref["soft"][:, :148], "soft")
    assert flag.all() and ((soft > 0.5) == bits.astype(bool)).all()


def test_wire_formats(dsp, oracle_best):
    """int16 {I,Q} in (unUSRPifyVector, radioInterface.cpp:91-116) and 148 soft bytes out ((char) round(soft*255.0),
    Transceiver.cpp:668-670): the wire-format entry points against the oracle fed the converted floats"""
    import torch
    rng = np.random.default_rng(0xC2F)
    nb = 936 * 2
    bits = np.stack([synth.normal_burst_bits(rng, 0) for _ in range(nb)])
    iq = np.zeros((500 * 864, 2), np.int16)
    dsp.tx_stream_host(bits, nb, iq)
    iq = np.clip(iq + np.rint(300.0 * rng.standard_normal(iq.shape)), -32768, 32767).astype(np.int16)   # ADC noise
    raw = (iq[:, 0].astype(np.float32) + 1j * iq[:, 1].astype(np.float32)).astype(np.complex64)
    tsc = np.zeros(nb, np.uint8)
    ref = oracle_best.rx_stream_demod(oracle_best.rx_resample_stream(raw, threads=4), nb, tsc, threads=8)
    ref_u8 = np.floor(ref["soft"][:, :148].astype(np.float64) * 255.0 + 0.5).astype(np.uint8)   # C round() for x >= 0
    for swap in (False, True):
        src = np.ascontiguousarray(iq[:, ::-1] if swap else iq)
        flag, amp, toa = np.zeros(nb, np.int32), np.zeros(nb, np.complex64), np.zeros(nb, np.float32)
        u8 = np.zeros((nb, 148), np.uint8)
        dsp.rx_stream_wire_host(src, 500, tsc, nb, flag, amp, toa, u8, swap_iq=swap)
        same(flag, ref["flag"], "flag"); same(amp, ref["amp"], "amp"); same(toa, ref["toa"], "toa")
        same(u8, ref_u8, "soft bytes (swap_iq=%s)" % swap)
    assert flag.all() and ((u8 > 127) == bits.astype(bool)).all()
    # device-level pieces
    dev = torch.device("cuda:0")
    d_iq = torch.from_numpy(iq.copy()).to(dev)
    res = torch.zeros(500 * 585 * 2, device=dev)
    dsp.resample_rx_i16_dev(d_iq, 500, res)
    same(res.cpu().numpy().view(np.complex64), oracle_best.rx_resample_stream(raw, threads=4), "int16 ingest resample")
    res2 = torch.zeros(400 * 585 * 2, device=dev)
    dsp.resample_rx_i16_dev(d_iq[100 * 864:], 400, res2, has_history=True)
    same(res2.cpu().numpy(), res.cpu().numpy()[100 * 585 * 2:], "int16 ingest with history")
    d_flag = torch.zeros(nb, dtype=torch.int32, device=dev); d_amp = torch.zeros(nb * 2, device=dev)
    d_toa = torch.zeros(nb, device=dev); d_u8 = torch.full((nb, 152), 7, dtype=torch.uint8, device=dev)
    dsp.demod_normal_u8_dev(res, 0, torch.zeros(nb, dtype=torch.uint8, device=dev), nb, d_flag, d_amp, d_toa, d_u8, 152)
    torch.cuda.synchronize()
    same(d_u8.cpu().numpy()[:, :148], ref_u8, "u8 rows at pitch 152")
    assert (d_u8.cpu().numpy()[:, 148:] == 0).all()


def test_full_size_properties(dsp):
    """BASELINE config-2 scale without the oracle: 10^5-frame-class stream, size-independent properties --
    TX->RX round trip recovers every bit, every burst detected, TOA on the 1/512 grid, and the result does not
    depend on how the stream is cut into launches (device-resident call vs segmented host pipeline)."""
    import torch
    dev = torch.device("cuda:0")
    nblocks = 160                                   # 160 * 936 = 149 760 bursts (4 680 warps: the wide-CTA detect launch), 40 000 chunks, 276 MB raw
    nb, nch = 936 * nblocks, 250 * nblocks
    g = torch.Generator(device=dev); g.manual_seed(2)
    bits = torch.randint(0, 2, (nb, 148), generator=g, device=dev, dtype=torch.uint8)
    bits[:, :3] = 0; bits[:, 145:] = 0
    mid = torch.from_numpy(synth.bits_of(synth.TSC[0]).copy()).to(dev)
    bits[:, 61:87] = mid
    iq = torch.zeros(nch * 864 * 2, dtype=torch.int16, device=dev)
    dsp.tx_stream_dev(bits, nb, iq)
    raw = iq.to(torch.float32) + 20.0 * torch.randn(iq.numel(), generator=g, device=dev)
    tsc = torch.zeros(nb, dtype=torch.uint8, device=dev)
    flag = torch.zeros(nb, dtype=torch.int32, device=dev); amp = torch.zeros(nb * 2, device=dev)
    toa = torch.zeros(nb, device=dev); soft = torch.zeros(nb * 148, device=dev)
    dsp.rx_stream_dev(raw, nch, tsc, nb, flag, amp, toa, soft, 148)
    torch.cuda.synchronize()
    hard = (soft.reshape(nb, 148) > 0.5).to(torch.uint8)
    assert bool(flag.all()) and bool((hard == bits).all())
    assert bool(((toa * 512) == torch.round(toa * 512)).all()) and float(toa.abs().max()) < 1.0
    raw_h = raw.cpu().numpy().view(np.complex64)
    f2, a2, t2 = np.zeros(nb, np.int32), np.zeros(nb, np.complex64), np.zeros(nb, np.float32)
    s2 = np.zeros((nb, 148), np.float32)
    dsp.rx_stream_host(raw_h, nch, tsc.cpu().numpy(), nb, f2, a2, t2, s2, 148)
    same(s2, soft.cpu().numpy().reshape(nb, 148), "host pipeline == device call")
    same(t2, toa.cpu().numpy(), "toa"); same(a2.view(np.float32), amp.cpu().numpy(), "amp")


def test_wide_detect_launch_equals_narrow_launches(dsp):
    """Batches of >= 4096 warps run detect as multi-warp CTAs whose warps meet at barriers (kernels.cu: BTS_DET_WARPS,
    BTS_DET_SYNC); smaller ones as one-warp CTAs, the form every oracle comparison above exercises.  Same bursts, both ways,
    with a ragged last CTA, a partial last warp, empty (gated) slots, mixed TSCs and channels: identical outputs."""
    import torch
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from tools import workloads
    dev = torch.device("cuda:0")
    bursts, tsc, _, occ = workloads.normal_batch(dsp, dev, n_arfcn=1025, frames=16, seed=77, empty=0.1)
    n = 4096 * 32 + 3 * 32 + 7                        # 4100 warps (820 five-warp CTAs), the last one 7 bursts wide
    assert n <= bursts.shape[0]

    def run(lo, hi, out):
        m = hi - lo
        flag, amp, toa, soft = out
        dsp.demod_normal_dev(bursts[lo:hi], 160, tsc[lo:hi], m, flag[lo:hi], amp[2 * lo:2 * hi], toa[lo:hi],
                             soft[lo * 148:hi * 148], 148, first=lo, gate_thr=5.0)
    def outs():
        return (torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n * 2, device=dev), torch.zeros(n, device=dev),
                torch.full((n * 148,), -1.0, device=dev))
    wide, narrow = outs(), outs()
    run(0, n, wide)
    step = 1024 * 32
    for lo in range(0, n, step):
        run(lo, min(n, lo + step), narrow)
    torch.cuda.synchronize()
    assert 0 < int(wide[0].sum()) < n                  # some slots detected, some gated or missed
    for a, b, what in zip(wide, narrow, ("flag", "amp", "toa", "soft")):
        same(a.cpu().numpy(), b.cpu().numpy(), what)


def test_cached_dfe_mode(dsp, oracle_best):
    """analyze -> designDFE -> equalize as three batched device calls (the cached-filter mode of
    Transceiver.cpp:315-396) gives the same result as the fused kernel"""
    import torch
    dev = torch.device("cuda:0")
    g = golden("normal_sps1.npz")
    n = g["bursts"].shape[0]
    bursts = torch.from_numpy(g["bursts"].view(np.float32).copy()).to(dev)
    lens = torch.from_numpy(g["lens"].copy()).to(dev); tsc = torch.from_numpy(g["tsc"].copy()).to(dev)
    flag = torch.zeros(n, dtype=torch.int32, device=dev); amp = torch.zeros(n * 2, device=dev)
    toa = torch.zeros(n, device=dev); chan = torch.zeros(n * 12, device=dev); off = torch.zeros(n, device=dev)
    dsp.analyze_dev(bursts, 160, tsc, n, flag, amp, toa, chan, off, lens=lens)
    torch.cuda.synchronize()
    a = amp.cpu().numpy().view(np.complex64); ok = flag.cpu().numpy() == 1
    same(flag.cpu().numpy(), g["flag"], "flag"); same(a, g["amp"], "amp"); same(toa.cpu().numpy(), g["toa"], "toa")
    same(off.cpu().numpy(), g["off"], "off")
    # caller glue (Transceiver.cpp:340-347, :391) on the host in float32, op for op
    f32 = np.float32
    ar, ai = a.real.astype(f32), a.imag.astype(f32)
    with np.errstate(all="ignore"):
        n2 = ai * ai + ar * ar
        ia_r, ia_i = ar / n2, (-ai) / n2                                   # complex(1,0)/amplitude
        snr = (n2.astype(np.float64) / (np.float64(f32(250.0) * f32(250.0)) + 1.0)).astype(f32)

    def cmul(xr, xi, sr, si):
        return xr * sr - xi * si, xr * si + xi * sr
    c_all = chan.cpu().numpy().view(np.complex64).reshape(n, 6)
    cr, ci = cmul(c_all.real.astype(f32), c_all.imag.astype(f32), ia_r[:, None], ia_i[:, None])
    chs = (cr + 1j * ci).astype(np.complex64)
    same(chs[ok], g["chan"][ok], "channel after 1/amp")
    br, bi = cmul(g["bursts"].real.astype(f32), g["bursts"].imag.astype(f32), ia_r[:, None], ia_i[:, None])
    scaled = np.where(ok[:, None], (br + 1j * bi).astype(np.complex64), 0).astype(np.complex64)
    snr = np.where(ok, snr, 1.0).astype(f32)
    chs = np.where(ok[:, None], chs, 0).astype(np.complex64); chs[~ok, 0] = 1
    w = torch.zeros(n * 14, device=dev); b = torch.zeros(n * 10, device=dev)
    dsp.design_dfe_dev(torch.from_numpy(chs.view(f32).copy()).to(dev), torch.from_numpy(snr).to(dev), n, w, b)
    soft = torch.zeros(n * 160, device=dev); after = torch.zeros(n * 160 * 2, device=dev)
    toa_eq = torch.from_numpy((g["toa"] - g["off"]).astype(f32)).to(dev)
    dsp.equalize_dev(torch.from_numpy(scaled.view(f32).copy()).to(dev), 160, n, toa_eq, w, b, soft, 160,
                     burst_out=after, out_pitch=160, lens=lens)
    torch.cuda.synchronize()
    same(w.cpu().numpy().view(np.complex64).reshape(n, 7)[ok], g["w"][ok], "w")
    same(b.cpu().numpy().view(np.complex64).reshape(n, 5)[ok], g["b"][ok], "b")
    same(soft.cpu().numpy().reshape(n, 160)[ok], g["soft"][ok], "soft via analyze + designDFE + equalize")
    # and the stand-alone equalizeBurst on one of them leaves the same delayed burst behind
    i = int(np.flatnonzero(ok)[0])
    s1, after1 = oracle_best.equalize(scaled[i, :g["lens"][i]], float(g["toa"][i] - g["off"][i]), g["w"][i], g["b"][i])
    same(after.cpu().numpy().view(np.complex64).reshape(n, 160)[i, :g["lens"][i]], after1, "burst after equalize")


@pytest.mark.gpu
def test_resampler_ragged_sizes(dsp, oracle_best):
    """RX resampler at sizes around every boundary of the tiled kernel (96 periods per step, one CTA per SM): a single
    chunk, partial last super-tiles, more super-tiles than SMs, with and without history, float and int16 input,
    output rows past the end untouched, and the unaligned-pointer fallback."""
    import torch
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(77)
    nmax = 1700                                       # 1700 chunks = 15300 periods = 160 super-tiles > 148 SMs
    iq = rng.integers(-3000, 3000, size=(nmax * 864, 2)).astype(np.int16)
    raw = (iq[:, 0].astype(np.float32) + 1j * iq[:, 1].astype(np.float32)).astype(np.complex64)
    want = oracle_best.rx_resample_stream(raw, threads=8)
    d_raw = torch.from_numpy(raw.view(np.float32).copy()).to(dev)
    d_iq = torch.from_numpy(iq.copy()).to(dev)
    for nch in (1, 2, 3, 10, 11, 32, 33, 107, 1579, nmax):
        for skip in (0, 5):                          # skip > 0: start inside the stream, with real history
            if skip + nch > nmax:
                continue
            res = torch.full(((nch + 1) * 585 * 2,), 7.5, device=dev)
            dsp.resample_rx_dev(d_raw[skip * 864 * 2:], nch, res, has_history=skip > 0)
            got = res.cpu().numpy()
            if skip == 0:
                same(got[:nch * 585 * 2].view(np.complex64), want[:nch * 585], "float nch=%d" % nch)
            else:
                # with history only the stream's very first chunk differs (zero history there), so compare from `skip`
                same(got[:nch * 585 * 2].view(np.complex64), want[skip * 585:(skip + nch) * 585], "float hist nch=%d" % nch)
            assert (got[nch * 585 * 2:] == 7.5).all(), "wrote past the end (nch=%d)" % nch
            res = torch.full(((nch + 1) * 585 * 2,), 7.5, device=dev)
            dsp.resample_rx_i16_dev(d_iq[skip * 864:], nch, res, has_history=skip > 0)
            got = res.cpu().numpy()
            same(got[:nch * 585 * 2].view(np.complex64), want[skip * 585:(skip + nch) * 585], "int16 nch=%d skip=%d" % (nch, skip))
            assert (got[nch * 585 * 2:] == 7.5).all()
    # several radios in one launch == one launch per radio (streams 5 chunks of slack apart, with and without history)
    S, nch = 7, 37
    pitch_in, pitch_out = (nch + 5) * 864, (nch + 5) * 585
    for hist in (False, True):
        base = 864 if hist else 0                     # with history, each stream starts one chunk into its slab
        many = torch.full((S * pitch_out * 2,), 7.5, device=dev)
        dsp.resample_rx_i16_streams_dev(d_iq[base:], pitch_in, S, nch, many, pitch_out, has_history=hist)
        got = many.cpu().numpy().reshape(S, pitch_out * 2)
        for a in range(S):
            one = torch.zeros(nch * 585 * 2, device=dev)
            dsp.resample_rx_i16_dev(d_iq[base + a * pitch_in:], nch, one, has_history=hist)
            same(got[a, :nch * 585 * 2], one.cpu().numpy(), "multi-stream %d hist=%s" % (a, hist))
            assert (got[a, nch * 585 * 2:] == 7.5).all()
    # an 8-byte-aligned (not 16) input pointer takes the one-CTA-per-chunk kernel: same results
    shifted = torch.zeros(40 * 864 * 2 + 2, device=dev)
    shifted[2:] = d_raw[:40 * 864 * 2]
    res = torch.zeros(40 * 585 * 2, device=dev)
    dsp.resample_rx_dev(shifted[2:], 40, res)
    same(res.cpu().numpy().view(np.complex64), want[:40 * 585], "unaligned fallback")


@pytest.mark.gpu
def test_tx_datagrams(dsp, oracle_best):
    """GSM-core -> transceiver datagrams through modulate / power scaling / slot placement / TX resample / int16
    (Transceiver.cpp:100-114, 582-632; radioInterface.cpp:123-168) against the reference functions under the same glue"""
    if oracle_best.kind != "ref":
        pytest.skip("needs the compiled reference")
    rng = np.random.default_rng(2024)
    nframes, fn0 = 234, 2715648 - 100                      # the window crosses the hyperframe wrap
    n = 1500
    dg = np.zeros((n, 156), np.uint8)                       # pitch 156 > 154
    fn = (fn0 + rng.integers(-5, nframes + 5, n)) % 2715648
    dg[:, 0] = rng.integers(0, 8, n)
    dg[::97, 0] = 9                                         # bad timeslot: dropped
    for k in range(4):
        dg[:, 1 + k] = (fn >> ((3 - k) * 8)) & 0xff
    dg[:, 5] = rng.choice(np.array([0, 3, 10, 19, 20, 37, 246], np.uint8), n)   # 246 = (char) -10: a gain of 10
    dg[:, 6:154] = rng.integers(0, 2, (n, 148))
    dummy = synth.bits_of("0001111101101110110000010100100111000001001000100000001111100011100010111000101110001010111010010100011001100111001111010011111000100101111101010000")
    for filler in (None, dummy):
        want, placed_ref = oracle_best.tx_datagrams(dg, fn0, nframes, filler)
        got, placed = dsp.tx_datagrams_host(dg, fn0, nframes, filler)
        assert placed == placed_ref and 0 < placed < n
        same(got, want, "TX datagrams (filler=%s)" % (filler is not None))
    assert np.abs(got).max() > 1000


@pytest.mark.gpu
def test_no_overrun_canaries(dsp, oracle_best):
    """device entry points must not write outside their outputs: sentinel-filled margins around every output buffer of
    the fused TX chain, the policy pull, the XCCH / RACH decoders and the soft-byte equaliser"""
    import torch
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    # fused TX chain: 936 bursts -> 250 chunks x 864 int16 pairs
    nb = 936
    bits = torch.from_numpy(rng.integers(0, 2, (nb, 148)).astype(np.uint8)).to(dev)
    iq = torch.full((250 * 864 * 2 + 64,), 12345, dtype=torch.int16, device=dev)
    dsp.tx_stream_dev(bits, nb, iq[32:])
    torch.cuda.synchronize()
    h = iq.cpu().numpy()
    assert (h[:32] == 12345).all() and (h[-32:] == 12345).all() and (h[32:-32] != 12345).any()
    # policy pull over pitched bursts: valid[n] and datagram rows at pitch 160
    A, F = 5, 7
    n = A * F * 8
    bursts = synth.make_trx_batch(oracle_best.modulate, oracle_best.expected_corr_type, F, [0] * A, np.ones((A, 8)), seed=6)
    d_b = torch.from_numpy(bursts.view(np.float32).copy()).to(dev)
    valid = torch.full((n + 16,), -7, dtype=torch.int32, device=dev)
    dg = torch.full(((n + 2) * 160,), 0xA5, dtype=torch.uint8, device=dev)
    trx = dsp.trx_create([0] * A, np.ones((A, 8), np.uint8), 0)
    dsp.trx_pull_dev(trx, d_b, 160, F, 0, valid[8:], dg[160:], 160)
    torch.cuda.synchronize()
    dsp.trx_destroy(trx)
    v, d = valid.cpu().numpy(), dg.cpu().numpy()
    assert (v[:8] == -7).all() and (v[-8:] == -7).all() and set(np.unique(v[8:-8])) <= {0, 1}
    assert (d[:160] == 0xA5).all() and (d[-160:] == 0xA5).all()
    # XCCH / RACH decoders
    nfr = 37
    soft = torch.from_numpy(rng.integers(0, 256, (nfr * 4, 148)).astype(np.uint8)).to(dev)
    u = torch.full((nfr * 228 + 64,), 9, dtype=torch.uint8, device=dev)
    ok = torch.full((nfr + 16,), -3, dtype=torch.int32, device=dev)
    dsp.xcch_decode_dev(soft, 148, nfr, u[32:], ok[8:])
    torch.cuda.synchronize()
    hu, hk = u.cpu().numpy(), ok.cpu().numpy()
    assert (hu[:32] == 9).all() and (hu[-32:] == 9).all() and (hu[32:-32] <= 1).all()
    assert (hk[:8] == -3).all() and (hk[-8:] == -3).all() and set(np.unique(hk[8:-8])) <= {0, 1}


@pytest.mark.gpu
def test_tx_streams_equal_single_stream_calls(dsp):
    """btsdsp_tx_streams_dev: many radios' TX chains in one launch == one btsdsp_tx_stream_dev call per radio (itself
    checked bit for bit against the reference's modulateBurst + pushBuffer in test_stream_kernels_match_golden), every
    stream from zero resampler history; nothing is written outside the output"""
    import torch
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11)
    for nb, ns in ((468, 5), (936, 3), (936 * 2, 151)):
        nch = nb // 4 * 625 // 585
        bits = torch.from_numpy(rng.integers(0, 2, (ns, nb, 148)).astype(np.uint8)).to(dev)
        multi = torch.full((ns * nch * 864 * 2 + 64,), 12345, dtype=torch.int16, device=dev)
        dsp.tx_streams_dev(bits, nb, ns, multi[32:])
        single = torch.zeros((ns, nch * 864 * 2), dtype=torch.int16, device=dev)
        for a in range(ns):
            dsp.tx_stream_dev(bits[a], nb, single[a])
        torch.cuda.synchronize()
        m = multi.cpu().numpy()
        assert (m[:32] == 12345).all() and (m[-32:] == 12345).all()
        same(m[32:-32].reshape(ns, -1), single.cpu().numpy(), "tx streams nb=%d ns=%d" % (nb, ns))


@pytest.mark.gpu
def test_rx_stream_in_pieces_equals_one_call(dsp):
    """btsdsp_rx_stream_cont_dev: a running stream handed over in 117-frame pieces, each seeing the 192 samples before it
    (RadioInterface::pullBuffer's rcvHistory), gives exactly the one-call result; without the history flag the first
    outputs of a later piece differ (the stream would restart from zeros)"""
    import torch
    dev = torch.device("cuda:0")
    nblocks = 6
    nb, nch = 936 * nblocks, 250 * nblocks
    g = torch.Generator(device=dev); g.manual_seed(4)
    bits = torch.randint(0, 2, (nb, 148), generator=g, device=dev, dtype=torch.uint8)
    bits[:, :3] = 0; bits[:, 145:] = 0
    bits[:, 61:87] = torch.from_numpy(synth.bits_of(synth.TSC[0]).copy()).to(dev)
    iq = torch.zeros(nch * 864 * 2, dtype=torch.int16, device=dev)
    dsp.tx_stream_dev(bits, nb, iq)
    raw = iq.to(torch.float32) + 300.0 * torch.randn(iq.numel(), generator=g, device=dev)
    tsc = torch.zeros(nb, dtype=torch.uint8, device=dev)

    def outs():
        return (torch.zeros(nb, dtype=torch.int32, device=dev), torch.zeros(nb * 2, device=dev), torch.zeros(nb, device=dev),
                torch.zeros(nb * 148, device=dev))
    f1, a1, t1, s1 = outs()
    dsp.rx_stream_dev(raw, nch, tsc, nb, f1, a1, t1, s1, 148)
    for hist in (True, False):
        f2, a2, t2, s2 = outs()
        for lo, hi in ((0, 1), (1, 4), (4, 6)):                       # pieces of 1, 3 and 2 blocks
            b0, n = lo * 936, (hi - lo) * 936
            dsp.rx_stream_cont_dev(raw.data_ptr() + lo * 250 * 864 * 8, hist and lo > 0, (hi - lo) * 250, tsc[b0:], n,
                                   f2[b0:], a2[2 * b0:], t2[b0:], s2[b0 * 148:], 148)
        torch.cuda.synchronize()
        equal = all(torch.equal(x, y) for x, y in ((f1, f2), (a1, a2), (t1, t2), (s1, s2)))
        assert equal == hist


@pytest.mark.gpu
def test_cuda_graph_replay_of_small_batches(dsp, oracle_best):
    """btsdsp_graph_*: three small demod calls recorded once and replayed give the directly-launched results (after the
    inputs change, too); a call that would move a scratch buffer a live graph refers to is refused"""
    import torch
    dev = torch.device("cuda:0")
    n = 96 * 3
    bursts, lens, tsc, _ = synth.make_normal_batch(oracle_best.modulate, n, seed=31)
    d_b = torch.from_numpy(bursts.view(np.float32).copy()).to(dev)
    d_t = torch.from_numpy(tsc).to(dev)
    d_l = torch.from_numpy(lens).to(dev)
    st = torch.cuda.Stream()

    def outs():
        return (torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n * 2, device=dev), torch.zeros(n, device=dev),
                torch.zeros(n * 148, device=dev))

    def calls(o, s):
        for k in range(3):
            lo = 96 * k
            dsp.demod_normal_dev(d_b[lo:], 160, d_t[lo:], 96, o[0][lo:], o[1][2 * lo:], o[2][lo:], o[3][lo * 148:], 148, lens=d_l[lo:],
                                 first=lo, stream=s)
    ref = outs()
    calls(ref, st)                               # direct launches (also sizes the stream's scratch before the capture)
    st.synchronize()
    got = outs()
    l0 = dsp.launch_count
    g = dsp.graph_begin(st)
    calls(got, st)
    dsp.graph_end(g, st)
    assert dsp.launch_count == l0                # recorded, not run
    st.synchronize()
    assert not got[0].any()
    dsp.graph_launch(g, st)
    st.synchronize()
    assert dsp.launch_count == l0 + 6
    assert all(torch.equal(a, b) for a, b in zip(got, ref))
    d_b[:96] = d_b[96:192].clone()               # same graph, new samples in the same buffers
    calls(ref, st)
    dsp.graph_launch(g, st)
    st.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(got, ref))
    big = 40000                                  # a call on that stream that needs more scratch than the graph's calls had
    with pytest.raises(Exception):
        dsp.demod_normal_dev(torch.zeros(big * 160 * 2, device=dev), 160, torch.zeros(big, dtype=torch.uint8, device=dev), big,
                             torch.zeros(big, dtype=torch.int32, device=dev), torch.zeros(big * 2, device=dev), torch.zeros(big, device=dev),
                             torch.zeros(big * 148, device=dev), 148, stream=st)
    dsp.graph_destroy(g)
    st.synchronize()


@pytest.mark.gpu
def test_resample_tx_with_history_and_unaligned(dsp):
    """btsdsp_resample_tx_dev (the tuned kernel with a loaded input tile): a later part of a stream, seeing the samples
    before it, equals the same chunks computed from the start; an output that is not 4-byte aligned is refused"""
    import torch
    g = golden("stream_sps1.npz")
    dev = torch.device("cuda:0")
    x = torch.from_numpy(g["stream_head"].view(np.float32).copy()).to(dev)
    want = g["iq_head"]
    full = torch.zeros(20 * 864 * 2, dtype=torch.int16, device=dev)
    dsp.resample_tx_dev(x, 20, full)
    same(full.cpu().numpy().reshape(-1, 2), want, "TX resample from the stream start")
    part = torch.zeros(15 * 864 * 2, dtype=torch.int16, device=dev)
    dsp.resample_tx_dev(x[5 * 585 * 2:], 15, part, has_history=True)
    same(part.cpu().numpy().reshape(-1, 2), want[5 * 864:], "TX resample, chunks 5..19 with history")
    cold = torch.zeros(15 * 864 * 2, dtype=torch.int16, device=dev)
    dsp.resample_tx_dev(x[5 * 585 * 2:], 15, cold, has_history=False)
    assert not np.array_equal(cold.cpu().numpy().reshape(-1, 2)[:8], want[5 * 864:5 * 864 + 8])   # zeros before the part
    same(cold.cpu().numpy().reshape(-1, 2)[864:], want[6 * 864:], "chunks after the first no longer see the cut")
    odd = torch.zeros(20 * 864 * 2 + 1, dtype=torch.int16, device=dev)
    with pytest.raises(Exception):                                        # {I,Q} pairs are written as 4-byte units
        dsp.resample_tx_dev(x, 20, odd[1:])


@pytest.mark.gpu
def test_copy_only_mode_moves_bytes_and_launches_nothing(dsp):
    """btsdsp_set_copy_only (bench.py's copy roofline leg): the host pipeline issues its copies and no kernel, results stay
    untouched; switching it off restores the normal call"""
    g = golden("stream_sps1.npz")
    raw = g["raw_head"]
    nch = raw.size // 864
    nb = (nch * 585 // 625) * 4
    tsc = np.zeros(nb, np.uint8)

    def run():
        f, a, t = np.full(nb, -5, np.int32), np.zeros(nb, np.complex64), np.zeros(nb, np.float32)
        s = np.full((nb, 148), -1.0, np.float32)
        dsp.rx_stream_host(raw, nch, tsc, nb, f, a, t, s, 148)
        return f, s
    f0, s0 = run()
    assert (f0 >= 0).all() and (s0 >= 0).all()
    dsp.set_copy_only(True)
    l0 = dsp.launch_count
    f1, s1 = run()
    assert dsp.launch_count == l0                    # not one kernel
    dsp.set_copy_only(False)
    f2, s2 = run()
    assert dsp.launch_count > l0
    same(f2, f0, "flags after copy-only mode"); same(s2, s0, "soft bits after copy-only mode")
