"""L1 FEC after the path: XCCH deinterleave + soft Viterbi + Fire-code check.  The checker is the
reference's own SoftVector::decode / ViterbiR2O4 / Parity classes under the restated ten lines of XCCHL1Decoder glue
(oracle/ref_shim.cpp); inputs are frames encoded by the reference's encoder, hit with noise / erasures / bursts of
errors and quantised to the RX datagram's soft bytes."""
import numpy as np
import pytest

import synth
from conftest import golden
from emu import Emu


def make_frames(enc, n, seed):
    """n L2 frames -> soft bytes (n*4, 148) with a spread of channel qualities; returns (soft, d)"""
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 2, (n, 184)).astype(np.uint8)
    e = enc(d)                                                   # (n*4, 114) e-bits
    soft = rng.integers(0, 256, (n * 4, 148)).astype(np.uint8)   # tails / midamble / stealing bits: irrelevant junk
    sigma = np.repeat(rng.choice([0.05, 0.2, 0.35, 0.5, 0.7, 1.0], n), 4)[:, None]
    x = (2.0 * e - 1.0) + sigma * rng.standard_normal(e.shape)   # BPSK + AWGN
    p = 1.0 / (1.0 + np.exp(-2.0 * x / np.maximum(sigma, 0.3) ** 2))
    b = np.clip(np.rint(p * 255.0), 0, 255).astype(np.uint8)
    lost = rng.random(n * 4) < 0.08                              # a whole burst missing: 0.5 everywhere (:627)
    b[lost] = 128
    sat = rng.random(b.shape) < 0.1                              # saturated decisions exercise the 0.01 clamps
    b[sat] = np.where(e[sat] > 0, 255, 0)
    wrong = rng.random(b.shape) < 0.01                           # confident errors
    b[wrong] = 255 - b[wrong]
    soft[:, 3:60] = b[:, :57]
    soft[:, 88:145] = b[:, 57:]
    return soft, d


@pytest.fixture(scope="module")
def frames(oracle_best):
    if oracle_best.kind != "ref":
        pytest.skip("frame generation uses the reference encoder")
    return make_frames(oracle_best.xcch_encode, 600, 21)


def check(got_u, got_ok, want_u, want_ok, d):
    assert np.array_equal(got_ok, want_ok)
    assert np.array_equal(got_u, want_u)
    good = want_ok.astype(bool)
    assert 0.3 < good.mean() < 0.98                              # both outcomes are exercised
    assert (want_u[good][:, :184] == d[good]).all()              # a clean syndrome means the payload is right


def test_fec_hostemu_matches_reference(oracle_best, hostemu, frames):
    soft, d = frames
    want_u, want_ok = oracle_best.xcch_decode(soft)
    got_u, got_ok = Emu(hostemu).xcch_decode(soft)
    check(got_u, got_ok, want_u, want_ok, d)
    # wider rows (the datagram pitch) and the golden fixture committed from the reference
    wide = np.zeros((soft.shape[0], 160), np.uint8)
    wide[:, 8:156] = soft
    u2, ok2 = Emu(hostemu).xcch_decode(np.ascontiguousarray(wide[:, 8:]))
    assert np.array_equal(u2[:, :], want_u) and np.array_equal(ok2, want_ok)


def test_fec_port_matches_reference_and_golden(oracle_port, oracle_best, frames):
    soft, d = frames
    want_u, want_ok = oracle_best.xcch_decode(soft)
    got_u, got_ok = oracle_port.xcch_decode(soft)
    check(got_u, got_ok, want_u, want_ok, d)
    g = golden("fec_sps1.npz")
    u, ok = oracle_port.xcch_decode(g["soft"])
    assert np.array_equal(u, g["u"]) and np.array_equal(ok, g["ok"])


def test_fec_hostemu_matches_golden(hostemu):
    g = golden("fec_sps1.npz")
    u, ok = Emu(hostemu).xcch_decode(g["soft"])
    assert np.array_equal(u, g["u"]) and np.array_equal(ok, g["ok"])


@pytest.mark.gpu
def test_fec_gpu_matches_reference(oracle_best, dsp, frames):
    soft, d = frames
    want_u, want_ok = oracle_best.xcch_decode(soft)
    got_u, got_ok = dsp.xcch_decode_host(soft)
    check(got_u, got_ok, want_u, want_ok, d)
    g = golden("fec_sps1.npz")
    u, ok = dsp.xcch_decode_host(g["soft"])
    assert np.array_equal(u, g["u"]) and np.array_equal(ok, g["ok"])
    one_u, one_ok = dsp.xcch_decode_host(soft[:4])               # a single frame
    assert np.array_equal(one_u, want_u[:1]) and one_ok[0] == want_ok[0]


def syndrome_frames(o):
    """clean frames whose parity word was XORed before encoding: syndrome = the XOR pattern.  The reference keeps the
    64-bit syndrome in an `unsigned` (GSML1FEC.cpp:652), so patterns confined to bits 32..39 still pass its check."""
    rng = np.random.default_rng(77)
    flips = np.array([0, 1 << 39, 1 << 32, 0xFF << 32, 1 << 31, 1, (1 << 39) | 1, 1 << 35], np.uint64)
    d = rng.integers(0, 2, (flips.size, 184)).astype(np.uint8)
    e = o.xcch_encode_pflip(d, flips)
    soft = np.full((flips.size * 4, 148), 128, np.uint8)
    b = np.where(e > 0, 250, 5).astype(np.uint8)
    soft[:, 3:60] = b[:, :57]
    soft[:, 88:145] = b[:, 57:]
    return soft, (flips & np.uint64(0xFFFFFFFF)) == 0


def test_syndrome_is_judged_on_its_low_32_bits(oracle_best, oracle_port, hostemu):
    if oracle_best.kind != "ref":
        pytest.skip("needs the reference classes")
    soft, passes = syndrome_frames(oracle_best)
    want_u, want_ok = oracle_best.xcch_decode(soft)
    assert np.array_equal(want_ok.astype(bool), passes)          # the reference accepts syndromes confined to bits 32..39
    for impl in (oracle_port, Emu(hostemu)):
        u, ok = impl.xcch_decode(soft)
        assert np.array_equal(u, want_u) and np.array_equal(ok, want_ok)


@pytest.mark.gpu
def test_syndrome_low_32_bits_gpu(oracle_best, dsp):
    soft, passes = syndrome_frames(oracle_best)
    want_u, want_ok = oracle_best.xcch_decode(soft)
    u, ok = dsp.xcch_decode_host(soft)
    assert np.array_equal(u, want_u) and np.array_equal(ok, want_ok) and np.array_equal(ok.astype(bool), passes)


def make_access(o, n, seed):
    """n access bursts' soft bytes: RA/BSIC encoded by the reference encoder, noise of four strengths"""
    rng = np.random.default_rng(seed)
    ra = rng.integers(0, 256, n).astype(np.uint8)
    bsic = rng.integers(0, 64, n).astype(np.uint8)
    c = o.rach_encode(ra, bsic)
    soft = rng.integers(0, 256, (n, 148)).astype(np.uint8)
    sig = rng.choice([0.1, 0.4, 0.8, 1.2], n)[:, None]
    x = (2.0 * c - 1) + sig * rng.standard_normal(c.shape)
    p = 1 / (1 + np.exp(-2 * x / np.maximum(sig, 0.3) ** 2))
    soft[:, 49:85] = np.clip(np.rint(p * 255), 0, 255)
    return soft, ra, bsic


def test_rach_decode_hostemu_matches_reference(oracle_best, hostemu):
    if oracle_best.kind != "ref":
        pytest.skip("needs the reference classes")
    soft, ra, bsic = make_access(oracle_best, 2000, 4)
    want = oracle_best.rach_decode(soft)
    got = Emu(hostemu).rach_decode(soft)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    valid = (want[1] == 0) & (want[2] == bsic)
    assert 0.5 < valid.mean() < 0.95 and (want[3][valid] == ra[valid]).all()


@pytest.mark.gpu
def test_rach_decode_gpu_matches_reference(oracle_best, dsp):
    if oracle_best.kind != "ref":
        pytest.skip("needs the reference classes")
    soft, ra, bsic = make_access(oracle_best, 2003, 4)           # not a multiple of the CTA's 8 warps
    want = oracle_best.rach_decode(soft)
    got = dsp.rach_decode_host(soft)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


@pytest.mark.gpu
def test_rach_chain_end_to_end(oracle_best, dsp):
    """RA / BSIC -> reference RACH encoder -> access burst -> GMSK -> delay, noise -> detectRACHBurst + demodulateBurst
    (GPU) -> datagram soft bytes -> RACH block decoder (GPU): every burst detected, tail clean, BSIC and RA recovered;
    and the decoder agrees with the reference classes on the same bytes"""
    if oracle_best.kind != "ref":
        pytest.skip("needs the reference encoder")
    rng = np.random.default_rng(12)
    n = 512
    ra = rng.integers(0, 256, n).astype(np.uint8)
    bsic = rng.integers(0, 64, n).astype(np.uint8)
    coded = oracle_best.rach_encode(ra, bsic)
    bursts = np.zeros((n, 160), np.complex64)
    lens = np.where(np.arange(n) % 4 == 0, 157, 156).astype(np.int32)
    for i in range(n):
        b = synth.access_burst_bits(rng)
        b[49:85] = coded[i]
        x = dsp.modulate(b, int(lens[i]) - 88)
        bursts[i, :lens[i]] = synth.impair(rng, x, amp=1500.0, delay=rng.integers(0, 30) + rng.random(), snr_db=18.0)
    r = dsp.rach_host(bursts, lens)
    assert r["flag"].all()
    soft_u8 = np.zeros((n, 148), np.uint8)
    soft_u8[:] = np.floor(r["soft"][:, :148].astype(np.float64) * 255.0 + 0.5).astype(np.uint8)   # Transceiver.cpp:669
    u, tail, b2, ra2 = dsp.rach_decode_host(soft_u8)
    assert (tail == 0).all() and np.array_equal(b2, bsic) and np.array_equal(ra2, ra)
    for g, w in zip((u, tail, b2, ra2), oracle_best.rach_decode(soft_u8)):
        assert np.array_equal(g, w)


@pytest.mark.gpu
def test_xcch_chain_end_to_end(oracle_best, dsp):
    """L2 frames -> reference XCCH encoder -> four normal bursts each -> GMSK -> 2-tap channel, delay, noise -> fused
    detect / DFE / equalise with soft-byte output (GPU) -> XCCH block decoder (GPU): every frame passes the Fire-code
    check with its payload intact"""
    if oracle_best.kind != "ref":
        pytest.skip("needs the reference encoder")
    import torch
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(31)
    nfr = 128
    d = rng.integers(0, 2, (nfr, 184)).astype(np.uint8)
    e = oracle_best.xcch_encode(d)                                   # (nfr*4, 114)
    n = nfr * 4
    bursts = np.zeros((n, 160), np.complex64)
    lens = np.where(np.arange(n) % 4 == 0, 157, 156).astype(np.int32)
    tsc = np.full(n, 2, np.uint8)
    for i in range(n):
        b = synth.normal_burst_bits(rng, 2)
        b[3:60] = e[i, :57]
        b[88:145] = e[i, 57:]
        x = dsp.modulate(b, int(lens[i]) - 148)
        bursts[i, :lens[i]] = synth.impair(rng, x, amp=2000.0, delay=rng.uniform(0, 2), chan2=0.35 * np.exp(1j * i), snr_db=16.0)
    d_b = torch.from_numpy(bursts.view(np.float32).copy()).to(dev)
    d_l = torch.from_numpy(lens).to(dev); d_t = torch.from_numpy(tsc).to(dev)
    flag = torch.zeros(n, dtype=torch.int32, device=dev); amp = torch.zeros(n * 2, device=dev); toa = torch.zeros(n, device=dev)
    u8 = torch.zeros((n, 148), dtype=torch.uint8, device=dev)
    dsp.demod_normal_u8_dev(d_b, 160, d_t, n, flag, amp, toa, u8, 148, lens=d_l)
    fu = torch.zeros(nfr * 228, dtype=torch.uint8, device=dev); fok = torch.zeros(nfr, dtype=torch.int32, device=dev)
    dsp.xcch_decode_dev(u8, 148, nfr, fu, fok)
    torch.cuda.synchronize()
    assert bool(flag.all()) and bool(fok.all())
    assert np.array_equal(fu.cpu().numpy().reshape(nfr, 228)[:, :184], d)


# ---- TCH/FACCH block decoder (GSML1FEC.cpp:1031-1210): 8-burst diagonal interleave, class-1 Viterbi + class-2 slice + parity /
#      tail checks for speech frames, the XCCH decode for stolen (FACCH) blocks ----
def make_tch_stream(o, nblocks, seed):
    """one traffic channel: nblocks blocks (a quarter of them stolen by FACCH), encoded by the reference's own encoder
    classes, through BPSK + AWGN of varying quality, erasures, saturated and wrong-confident soft bytes"""
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 2, (nblocks, 260)).astype(np.uint8)
    f = rng.integers(0, 2, (nblocks, 184)).astype(np.uint8)
    steal = (rng.random(nblocks) < 0.25).astype(np.uint8)
    bits = o.tch_encode(d, f, steal).astype(np.float64)                # (4n+4, 148)
    nb = bits.shape[0]
    sigma = np.repeat(rng.choice([0.05, 0.2, 0.35, 0.5, 0.7, 1.0], (nb + 7) // 8), 8)[:nb, None]
    x = (2.0 * bits - 1.0) + sigma * rng.standard_normal(bits.shape)
    p = 1.0 / (1.0 + np.exp(-2.0 * x / np.maximum(sigma, 0.3) ** 2))
    soft = np.clip(np.rint(p * 255.0), 0, 255).astype(np.uint8)
    soft[rng.random(nb) < 0.05] = 128                                  # a burst lost entirely
    sat = rng.random(soft.shape) < 0.1
    soft[sat] = np.where(bits[sat] > 0, 255, 0)
    wrong = rng.random(soft.shape) < 0.01
    soft[wrong] = 255 - soft[wrong]
    return soft, d, f, steal


def check_tch(got, want, d, f, steal):
    for k in ("stolen", "good", "fok", "d", "fu"):
        assert np.array_equal(got[k], want[k]), k
    st = want["stolen"].astype(bool)
    assert 0.1 < st.mean() < 0.45 and (st == steal.astype(bool)).mean() > 0.9          # the flag itself rides a noisy bit
    g = want["good"].astype(bool)
    assert 0.3 < g[~st].mean() < 0.99 and not g[st].any()
    ok_speech = g & ~steal.astype(bool)
    # a good speech frame carries the right class-1A bits (what its parity protects); most are right everywhere in class 1
    assert (want["d"][ok_speech][:, :50] == d[ok_speech][:, :50]).mean() > 0.995
    fk = want["fok"].astype(bool) & steal.astype(bool)
    assert fk.any() and (want["fu"][fk][:, :184] == f[fk]).all()


@pytest.fixture(scope="module")
def tch_stream(oracle_best):
    if oracle_best.kind != "ref":
        pytest.skip("stream generation uses the reference encoder")
    return make_tch_stream(oracle_best, 500, 33)


def test_tch_hostemu_matches_reference(oracle_best, hostemu, tch_stream):
    soft, d, f, steal = tch_stream
    want = oracle_best.tch_decode(soft)
    check_tch(Emu(hostemu).tch_decode(soft), want, d, f, steal)
    wide = np.zeros((soft.shape[0], 160), np.uint8)                    # the RX datagram's row pitch
    wide[:, 8:156] = soft
    got = Emu(hostemu).tch_decode(np.ascontiguousarray(wide[:, 8:]))
    assert all(np.array_equal(got[k], want[k]) for k in want)


def test_tch_port_matches_reference(oracle_port, oracle_best, tch_stream):
    soft, d, f, steal = tch_stream
    want, got = oracle_best.tch_decode(soft), oracle_port.tch_decode(soft)
    assert all(np.array_equal(got[k], want[k]) for k in want)


def test_tch_clean_channel_round_trip(oracle_best, hostemu):
    """noise-free: every speech frame comes back whole and good, every FACCH payload too"""
    if oracle_best.kind != "ref":
        pytest.skip("needs the reference encoder")
    rng = np.random.default_rng(5)
    n = 40
    d = rng.integers(0, 2, (n, 260)).astype(np.uint8)
    f = rng.integers(0, 2, (n, 184)).astype(np.uint8)
    steal = (np.arange(n) % 5 == 2).astype(np.uint8)
    soft = (oracle_best.tch_encode(d, f, steal) * 255).astype(np.uint8)
    for r in (oracle_best.tch_decode(soft), Emu(hostemu).tch_decode(soft)):
        st = steal.astype(bool)
        assert np.array_equal(r["stolen"], steal) and r["good"][~st].all() and r["fok"][st].all()
        assert np.array_equal(r["d"][~st], d[~st]) and np.array_equal(r["fu"][st][:, :184], f[st])


@pytest.mark.gpu
def test_tch_gpu_matches_reference(oracle_best, dsp, tch_stream):
    import torch
    soft, d, f, steal = tch_stream
    want = oracle_best.tch_decode(soft)
    check_tch(dsp.tch_decode_host(soft), want, d, f, steal)
    # device entry point on datagram-pitched rows, outputs behind canaries
    dev = torch.device("cuda:0")
    n = soft.shape[0] // 4 - 1
    wide = torch.zeros((soft.shape[0], 160), dtype=torch.uint8, device=dev)
    wide[:, 8:156] = torch.from_numpy(soft).to(dev)
    dd = torch.full((n * 260 + 64,), 9, dtype=torch.uint8, device=dev)
    fu = torch.full((n * 228 + 64,), 9, dtype=torch.uint8, device=dev)
    gi = torch.full((3, n + 16), -3, dtype=torch.int32, device=dev)
    dsp.tch_decode_dev(wide.data_ptr() + 8, 160, n, dd[32:], gi[0, 8:], gi[1, 8:], fu[32:], gi[2, 8:])
    torch.cuda.synchronize()
    hd, hf, hg = dd.cpu().numpy(), fu.cpu().numpy(), gi.cpu().numpy()
    assert (hd[:32] == 9).all() and (hd[-32:] == 9).all() and (hf[:32] == 9).all() and (hf[-32:] == 9).all()
    assert (hg[:, :8] == -3).all() and (hg[:, -8:] == -3).all()
    assert np.array_equal(hd[32:-32].reshape(n, 260), want["d"]) and np.array_equal(hf[32:-32].reshape(n, 228), want["fu"])
    assert np.array_equal(hg[0, 8:-8], want["good"]) and np.array_equal(hg[1, 8:-8], want["stolen"]) and np.array_equal(hg[2, 8:-8], want["fok"])


def test_tch_port_and_hostemu_match_golden(oracle_port, hostemu):
    """fixture written by the compiled reference (oracle/gen_golden_r3.py): holds on a box without the reference too"""
    g = golden("tch_sps1.npz")
    for r in (oracle_port.tch_decode(g["soft"]), Emu(hostemu).tch_decode(g["soft"])):
        for k in ("d", "good", "stolen", "fu", "fok"):
            assert np.array_equal(r[k], g[k]), k
    assert 0 < g["stolen"].sum() < g["stolen"].size and g["good"].any()
