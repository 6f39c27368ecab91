#!/usr/bin/env python3
"""Golden vectors for the transmit-side L1 encoders, made with the COMPILED REFERENCE (oracle/_ref; run in the container that
has /root/reference): tests/golden/fec_encode.npz -- L2 frames / speech frames and the bursts XCCHL1Encoder / TCHFACCHL1Encoder
(the reference's classes under the restated glue, oracle/ref_shim.cpp) make of them.  The committed fixture lets a box without
the reference check the host emulation of the kernels, and the GPU test checks the kernels against it as well.

    python oracle/gen_golden_r4.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.oracle import Oracle  # noqa: E402
import test_fec_encode as te  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    R = Oracle("ref", sps=1)
    frames = te.make_xcch(40, 1234)
    xb = R.xcch_send_frames(frames, True, 2)
    d, f, steal = te.make_tch(48, 4321)
    tb = te.ref_tch(R, d, f, steal, True, 5, [(0, 48)])
    fn = os.path.join(OUT, "fec_encode.npz")
    np.savez_compressed(fn, frames=frames, xcch_bursts=xb, d=d, f=f, steal=steal, tch_bursts=tb)
    print(fn, os.path.getsize(fn))


if __name__ == "__main__":
    main()
