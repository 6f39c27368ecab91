#!/usr/bin/env python3
"""Golden vectors added in round 3, made with the COMPILED REFERENCE (oracle/_ref; run in the container that has
/root/reference): tests/golden/tch_sps1.npz -- a traffic channel's soft bytes and what TCHFACCHL1Decoder makes of them;
tests/golden/trx52_sps1.npz -- the second transceiver variant's receive policy (Transceiver52M) over a batch, for
mMaxExpectedDelay 1 (no equaliser) and 4 (windowed search + DFE).  The committed fixtures let a box without the reference
check the plain-C port and the host emulation of the kernels.

    python oracle/gen_golden_r3.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.oracle import Oracle, Oracle52  # noqa: E402
import test_fec  # noqa: E402
import test_variant52m as tv  # noqa: E402
import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    R = Oracle("ref", sps=1)
    soft, d, f, steal = test_fec.make_tch_stream(R, 64, 7)
    r = R.tch_decode(soft)
    np.savez_compressed(os.path.join(OUT, "tch_sps1.npz"), soft=soft, **r)
    o52 = Oracle52(1)
    bursts = synth.make_trx_batch(R.modulate, R.expected_corr_type, tv.NFRAMES, tv.TSC, tv.CHAN_TYPES, fn0=tv.FN0, seed=52)
    out = {"sha1": np.frombuffer(hashlib.sha1(bursts.tobytes()).digest(), np.uint8)}
    for md in (1, 4):
        v, dg, st = tv.oracle52_pull(o52, bursts, md, [(0, 29), (29, 30), (30, tv.NFRAMES)])
        out["valid%d" % md] = v
        out["dgram%d" % md] = dg
        out["state%d" % md] = st.view(np.uint8).reshape(len(tv.TSC), -1)
    np.savez_compressed(os.path.join(OUT, "trx52_sps1.npz"), **out)
    for fn in ("tch_sps1.npz", "trx52_sps1.npz"):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
