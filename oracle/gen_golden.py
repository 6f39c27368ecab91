#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the compiled reference (oracle/_ref/libref_oracle.so).

TEST TOOLING.  The reference has no golden vectors for this path, so the known answers
committed under tests/golden/ are outputs of the UNMODIFIED reference sources compiled in place
(oracle/Makefile target `ref`) on seeded synthetic inputs (tests/synth.py).  Inputs and outputs are both
stored, so the fixtures check the plain-C port, the host emulation and the CUDA path anywhere -- also on
the GPU box, where /root/reference does not exist.

    python oracle/gen_golden.py            # run from the repo root, in the build container
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.oracle import Oracle, NO_DELAY, FULL_SPAN  # noqa: E402
import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def tables(R, sps):
    d = {}
    for t, name in enumerate(["cos", "sin", "rot", "revrot", "pulse"]):
        d[name] = R.table(t)
    d["mid_seq"] = np.stack([R.table(5, i) for i in range(8)])
    d["mid_meta"] = np.stack([R.table(6, i) for i in range(8)])
    d["rach_seq"] = R.table(7)
    d["rach_meta"] = R.table(8)
    lr = R.table(9).copy()
    lr[960] = 0.0          # the reference reads one float past its 960-entry table
    d["lpf_rx"] = lr
    d["lpf_tx"] = R.table(10)
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    R = Oracle("ref", sps=1)
    assert abs(float(R.table(11)[0])) < 1e-30
    mod = lambda b, g: R.modulate(b, g)  # noqa: E731

    np.savez_compressed(os.path.join(OUT, "tables_sps1.npz"), **tables(R, 1))

    # config 1 (Transceiver/sigProcLibTest.cpp flow at sps 1): known bits, TSC 0, delay 6.932 applied by
    # the reference's own delayVector, 2-tap channel [9000, 3600], seeded AWGN
    bits = synth.bits_of("0000101010100111110010101010010110101110011000111001101010000" + synth.TSC[0] +
                         "0000101010100111110010101010010110101110011000111001101010000")
    tx = R.modulate(bits, 8)
    dl = R.delay_vector(tx, 6.932)
    ch = R.convolve(dl, np.array([9000, 3600, 0, 0], np.complex64), NO_DELAY)
    rng = np.random.default_rng(0xB2000001)
    rx = (ch + np.sqrt(0.001 / np.sqrt(2) / 2) * (rng.standard_normal(ch.size) + 1j * rng.standard_normal(ch.size))
          ).astype(np.complex64)
    ok, amp, toa, chan, off = R.analyze(rx, 0, 8.0, request=True)
    soft_slicer = R.demodulate(rx, amp, toa)
    chs = R.scale_vector(chan, 1.0 / complex(amp)) if False else chan
    w, b = R.design_dfe(chan, 1.0 / 0.001, 7)
    soft_dfe, rx_after = R.equalize(rx, toa - off, w, b)
    pk, pidx, pavg = R.peak_detect(R.correlate(rx[56:92], R.table(5, 0), NO_DELAY))
    np.savez_compressed(os.path.join(OUT, "config1_sps1.npz"), bits=bits, tx=tx, delayed=dl, chan_out=ch, rx=rx,
                        ok=ok, amp=amp, toa=toa, chan=chan, off=off, soft_slicer=soft_slicer, w=w, b=b,
                        soft_dfe=soft_dfe, rx_after=rx_after, pk=pk, pidx=pidx, pavg=pavg,
                        full=R.convolve(rx, w, FULL_SPAN))

    # normal bursts, config-4 style (mixed TSC 0-7 -> SURVEY F6), some noise-only
    bursts, lens, tsc, nbits = synth.make_normal_batch(mod, 96, seed=0xB2000004, noise_only=0.08)
    r = R.rx_normal_batch(bursts, lens, tsc)
    np.savez_compressed(os.path.join(OUT, "normal_sps1.npz"), bursts=bursts, lens=lens, tsc=tsc, bits=nbits, **r)

    # access bursts, config-3 style
    rb, rl, rbits, rdel = synth.make_rach_batch(mod, 96, seed=0xB2000003)
    r = R.rx_rach_batch(rb, rl)
    np.savez_compressed(os.path.join(OUT, "rach_sps1.npz"), bursts=rb, lens=rl, bits=rbits, delays=rdel, **r)

    # streams, config-2/5 style: 936 bursts = 117 frames = 250 chunks; keep 936 bursts' bits, first 20 chunks of TX
    rng = np.random.default_rng(0xB2000002)
    nb = 936
    sb = np.stack([synth.normal_burst_bits(rng, 0) for _ in range(nb)])
    stream = R.modulate_stream(sb)
    assert stream.size == 250 * 585
    iq = R.tx_resample_stream(stream)
    raw = (iq[:, 0] + 1j * iq[:, 1]).astype(np.complex64)
    raw = (raw + 50.0 * (rng.standard_normal(raw.size) + 1j * rng.standard_normal(raw.size))).astype(np.complex64)
    res = R.rx_resample_stream(raw)
    d = R.rx_stream_demod(res, nb, np.zeros(nb, np.uint8))
    keep = 24                                 # chunks of raw kept in the fixture (inputs are the big part)
    np.savez_compressed(os.path.join(OUT, "stream_sps1.npz"), bits=sb[:64], stream_head=stream[:20 * 585],
                        iq_head=iq[:20 * 864], raw_head=raw[:keep * 864], res_head=res[:keep * 585],
                        flag=d["flag"][:64], amp=d["amp"][:64], toa=d["toa"][:64], soft=d["soft"][:64],
                        ber=np.mean((d["soft"][:, :148] > 0.5) != sb))

    # sps = 4: modulate / delay / analyze / demodulateBurst only
    R.setup(4)
    np.savez_compressed(os.path.join(OUT, "tables_sps4.npz"), **tables(R, 4))
    rng = np.random.default_rng(0xB2000044)
    xs, outs = [], dict(ok=[], amp=[], toa=[], chan=[], off=[], soft=[], tx=[])
    tscs = []
    for i in range(16):
        t = i % 8
        b = synth.normal_burst_bits(rng, t)
        tx = R.modulate(b, 8)
        rx = R.delay_vector(tx * np.complex64(2000.0 * np.exp(1j * rng.uniform(0, 6.28))), rng.uniform(0, 8.0))
        rx = (rx + 40.0 * (rng.standard_normal(rx.size) + 1j * rng.standard_normal(rx.size))).astype(np.complex64)
        ok, amp, toa, chan, off = R.analyze(rx, t, 3.0, request=True)
        soft = R.demodulate(rx, amp if ok else 1.0, toa)
        xs.append(rx); tscs.append(t)
        for k, v in zip(("ok", "amp", "toa", "chan", "off", "soft", "tx"), (ok, amp, toa, chan, off, soft, tx)):
            outs[k].append(v)
    np.savez_compressed(os.path.join(OUT, "sps4.npz"), rx=np.stack(xs), tsc=np.array(tscs, np.uint8),
                        **{k: np.stack(v) for k, v in outs.items()})
    R.setup(1)
    # ---- caller policy (pullRadioVector + driveReceiveFIFO): 120 frames x 3 ARFCN, in three pulls.  The 2880 input
    #      bursts (3.7 MB) are NOT stored: tests regenerate them from the seed with tests/synth.make_trx_batch and check
    #      the stored SHA-1 first, so a fixture mismatch cannot be mistaken for a parity failure.
    import hashlib
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_trx_policy as tp
    nframes, fn0 = tp.GOLDEN_SHAPE
    bursts = synth.make_trx_batch(R.modulate, R.expected_corr_type, nframes, tp.TSC, tp.CHAN_TYPES, fn0=fn0)
    v, d, st = tp.oracle_pull(R, bursts, nframes, fn0, fn0 - 3, tp.GOLDEN_SPLIT)
    np.savez_compressed(os.path.join(OUT, "trx_sps1.npz"), sha1=np.frombuffer(hashlib.sha1(bursts.tobytes()).digest(), np.uint8),
                        valid=v, dgram=d, state=st.view(np.uint8).reshape(len(tp.TSC), -1))
    # ---- L1 FEC (XCCH block decoder): 96 frames encoded by the reference, impaired, decoded by the reference
    import test_fec
    soft, d = test_fec.make_frames(R.xcch_encode, 96, 5)
    u, ok = R.xcch_decode(soft)
    np.savez_compressed(os.path.join(OUT, "fec_sps1.npz"), soft=soft, d=d, u=u, ok=ok)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
