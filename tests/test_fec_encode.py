"""L1 encoders on the transmit side -- the producers of the 148-bit bursts modulateBurst is called with: XCCHL1Encoder
(sendFrame / encode / interleave / transmit, GSML1FEC.cpp:763-850) and TCHFACCHL1Encoder (encodeTCH / dispatch / interleave,
:1248-1392).  The checker runs the reference's own flow with its own classes (BitVector::LSB8MSB, Parity, BitVector::encode,
gTrainingSequence) under the restated glue (oracle/ref_shim.cpp: ref_xcch_send_frames, ref_tch_dispatch -- the latter keeps the
encoder's members mI[], mOffset, mPreviousFACCH exactly as dispatch() does, where the product writes block q straight into
bursts 4q .. 4q+7).  Bit-exact: these are bits."""
import numpy as np
import pytest

from conftest import golden
from emu import Emu


def make_xcch(n, seed):
    rng = np.random.default_rng(seed)
    f = rng.integers(0, 2, (n, 184)).astype(np.uint8)
    f[0] = 0                                                    # all-zero and all-one frames: the parity word's extremes
    f[1] = 1
    return f


def make_tch(n, seed, p_steal=0.3):
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 2, (n, 260)).astype(np.uint8)
    f = rng.integers(0, 2, (n, 184)).astype(np.uint8)
    steal = (rng.random(n) < p_steal).astype(np.uint8)
    steal[:4] = [0, 1, 1, 0]                                    # every (previous, current) pair of stealing flags
    return d, f, steal


def ref_tch(o, d, f, steal, lsb, tsc, splits):
    """the reference encoder over the blocks, one or several calls (its members carried between them)"""
    out, state = [], None
    for lo, hi in splits:
        b, state = o.tch_dispatch(d[lo:hi], f[lo:hi], steal[lo:hi], lsb, tsc, state)
        out.append(b)
    return np.concatenate(out)


def our_tch(enc, d, f, steal, lsb, tsc, splits, pitch=148):
    """the product (or its host emulation) the same way: the four closing bursts of a call are the next call's carry"""
    out, carry = [], None
    for lo, hi in splits:
        b = enc(d[lo:hi], f[lo:hi], steal[lo:hi], lsb, tsc, carry, pitch)
        out.append(b[:-4])
        carry = b[-4:]
    return np.concatenate(out), carry


@pytest.fixture(scope="module")
def ref(oracle_best):
    if oracle_best.kind != "ref":
        pytest.skip("needs the compiled reference")
    return oracle_best


@pytest.mark.parametrize("lsb,tsc", [(True, 2), (False, -1), (True, 7)])
def test_xcch_encode_hostemu_matches_reference(ref, hostemu, lsb, tsc):
    f = make_xcch(300, 31)
    want = ref.xcch_send_frames(f, lsb, tsc)
    got = Emu(hostemu).xcch_encode(f, lsb, tsc)
    assert np.array_equal(got, want)
    assert want[:, 60].all() and want[:, 87].all() and not want[:, :3].any() and not want[:, 145:].any()
    wide = Emu(hostemu).xcch_encode(f, lsb, tsc, burst_pitch=160)
    assert np.array_equal(wide[:, :148], want) and not wide[:, 148:].any()


@pytest.mark.parametrize("lsb,tsc", [(True, 2), (False, -1), (False, 5)])
def test_xcch_encode_lane_form_matches_reference(ref, hostemu, lsb, tsc):
    """the kernels' lane form (bit-packed words, compile-time bit positions, byte-wise CRC table) on the CPU"""
    f = make_xcch(300, 32)
    assert np.array_equal(Emu(hostemu).xcch_encode_lanes(f, lsb, tsc), ref.xcch_send_frames(f, lsb, tsc))
    assert np.array_equal(Emu(hostemu).xcch_encode_lanes(f, lsb, tsc, popc=True), ref.xcch_send_frames(f, lsb, tsc))   # parity by masks


def test_xcch_encode_agrees_with_the_decoder_test_generator(ref, hostemu):
    """the e-bits are the ones the (older) decoder-side generator ref_xcch_encode makes from the same d"""
    f = make_xcch(50, 5)
    e = ref.xcch_encode(f)
    got = Emu(hostemu).xcch_encode(f, False, -1)
    assert np.array_equal(got[:, 3:60], e[:, :57]) and np.array_equal(got[:, 88:145], e[:, 57:])


@pytest.mark.parametrize("splits", [[(0, 96)], [(0, 1), (1, 2), (2, 50), (50, 96)]])
def test_tch_encode_hostemu_matches_reference(ref, hostemu, splits):
    d, f, steal = make_tch(96, 77)
    for lsb, tsc in ((True, 5), (False, -1)):
        want = ref_tch(ref, d, f, steal, lsb, tsc, splits)
        got, carry = our_tch(Emu(hostemu).tch_encode, d, f, steal, lsb, tsc, splits)
        assert np.array_equal(got, want)
        # the closing bursts hold only the last block's half: odd e-bits and Hl
        assert not carry[:, 87].any() and (carry[:, 60] == steal[-1]).all()


def test_tch_encode_lane_form_matches_reference(ref, hostemu):
    """the traffic-channel kernels' lane form on the CPU: every group but the first (which needs the carry) against the reference"""
    d, f, steal = make_tch(96, 80)
    for lsb, tsc in ((True, 5), (False, -1)):
        want = ref_tch(ref, d, f, steal, lsb, tsc, [(0, 96)])
        seq = Emu(hostemu).tch_encode(d, f, steal, lsb, tsc)
        got = Emu(hostemu).tch_encode_lanes(d, f, steal, lsb, tsc)
        assert np.array_equal(got[4:-4], want[4:])
        assert np.array_equal(got[-4:], seq[-4:])                        # the closing (carry) group


def test_encode_decode_round_trip(hostemu):
    """encoder -> ideal soft bytes -> the receive-side block decoders give the frames back (no reference needed)"""
    emu = Emu(hostemu)
    f = make_xcch(40, 9)
    b = emu.xcch_encode(f, False, 3)
    u, ok = emu.xcch_decode(np.where(b > 0, 255, 0).astype(np.uint8))
    assert ok.all() and np.array_equal(u[:, :184], f)
    d, ff, steal = make_tch(40, 10)
    b = emu.tch_encode(d, ff, steal, False, 3)
    r = emu.tch_decode(np.where(b > 0, 255, 0).astype(np.uint8))
    sp = steal == 0
    assert np.array_equal(r["stolen"] != 0, ~sp)
    assert r["good"][sp].all() and np.array_equal(r["d"][sp], d[sp])
    assert r["fok"][~sp].all() and np.array_equal(r["fu"][~sp][:, :184], ff[~sp])


def test_encode_hostemu_matches_golden(hostemu):
    g = golden("fec_encode.npz")
    emu = Emu(hostemu)
    assert np.array_equal(emu.xcch_encode(g["frames"], True, 2), g["xcch_bursts"])
    assert np.array_equal(emu.xcch_encode_lanes(g["frames"], True, 2), g["xcch_bursts"])
    got, _ = our_tch(emu.tch_encode, g["d"], g["f"], g["steal"], True, 5, [(0, 10), (10, 48)])
    assert np.array_equal(got, g["tch_bursts"])


@pytest.mark.gpu
def test_xcch_encode_gpu_matches_reference(ref, dsp):
    f = make_xcch(5000, 41)
    for lsb, tsc, pitch in ((True, 2, 148), (False, -1, 148), (True, 0, 160)):
        want = ref.xcch_send_frames(f, lsb, tsc)
        got = dsp.xcch_encode_host(f, lsb, tsc, pitch)
        assert np.array_equal(got[:, :148], want)
        assert not got[:, 148:].any()
    g = golden("fec_encode.npz")
    assert np.array_equal(dsp.xcch_encode_host(g["frames"], True, 2), g["xcch_bursts"])


@pytest.mark.gpu
def test_tch_encode_gpu_matches_reference(ref, dsp):
    d, f, steal = make_tch(4000, 78)
    for splits in ([(0, 4000)], [(0, 1), (1, 3), (3, 1000), (1000, 4000)]):
        for lsb, tsc, pitch in ((True, 5, 148), (False, -1, 152)):
            want = ref_tch(ref, d, f, steal, lsb, tsc, splits)
            got, carry = our_tch(dsp.tch_encode_host, d, f, steal, lsb, tsc, splits, pitch)
            assert np.array_equal(got[:, :148], want)
            assert not carry[:, 87].any() and (carry[:, 60] == steal[-1]).all()


@pytest.mark.gpu
def test_tch_encode_edges_gpu(ref, dsp):
    """no blocks at all (the carry is completed with an empty block), one block, and a CTA boundary (8 groups per CTA)"""
    import torch
    d, f, steal = make_tch(17, 79)
    want = ref_tch(ref, d, f, steal, True, 5, [(0, 17)])
    first = dsp.tch_encode_host(d[:9], f[:9], steal[:9], True, 5)            # 10 groups: two CTAs
    assert np.array_equal(first[:-4], want[:36])
    dev = torch.device("cuda:0")
    carry = torch.from_numpy(first[-4:].copy()).to(dev)
    out = torch.full((4, 148), 7, dtype=torch.uint8, device=dev)
    z = torch.zeros((1, 260), dtype=torch.uint8, device=dev)
    dsp.tch_encode_dev(z, z, z, 0, 1, 5, carry, out, 148)
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    # closing the carry with nothing: the odd half and Hl of the carry, midamble, everything else zero
    exp = first[-4:].copy()
    assert np.array_equal(o, exp)
    rest = dsp.tch_encode_host(d[9:], f[9:], steal[9:], True, 5, carry=first[-4:])
    assert np.array_equal(rest[:-4], want[36:])
    one = dsp.tch_encode_host(d[:1], f[:1], steal[:1], False, -1)
    assert np.array_equal(one[:4], ref_tch(ref, d[:1], f[:1], steal[:1], False, -1, [(0, 1)]))


@pytest.mark.gpu
def test_encode_modulate_demod_decode_chain(ref, dsp):
    """L2 frames -> GPU encoder -> GPU modulator -> (clean channel) GPU normal-burst demod -> GPU block decoder: the frames
    come back, and the bursts in the middle are the reference encoder's"""
    f = make_xcch(64, 12)
    tsc = 2
    bits = dsp.xcch_encode_host(f, True, tsc)
    assert np.array_equal(bits, ref.xcch_send_frames(f, True, tsc))
    n = bits.shape[0]
    bursts = np.zeros((n, 160), np.complex64)
    for i in range(n):
        bursts[i, :156] = ref.modulate(bits[i], 8) * 1000.0
    lens = np.full(n, 156, np.int32)
    out = dsp.demod_normal_host(bursts, lens, np.full(n, tsc, np.uint8))
    assert out["flag"].all()
    soft = np.clip(np.rint(out["soft"][:, :148] * 255.0), 0, 255).astype(np.uint8)
    u, ok = dsp.xcch_decode_host(soft)
    lsb = np.arange(184).reshape(23, 8)[:, ::-1].reshape(-1)             # LSB8MSB as an index map
    assert ok.all() and np.array_equal(u[:, :184], f[:, lsb])
