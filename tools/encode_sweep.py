#!/usr/bin/env python3
"""One-off wide parity sweep of the transmit-side encoders on the GPU box: many random frames / blocks through the C ABI
against the compiled reference's own encoder flow (oracle ref_xcch_send_frames / ref_tch_dispatch), in one call and in
several with the carry.  Measurement aid."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import openbts_ttsou_b200 as pkg
from oracle.oracle import Oracle
import test_fec_encode as te

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
o = Oracle("ref")
dsp = pkg.BtsDsp(0, 1)
ok = True
for lsb, tsc in ((True, 0), (False, 7), (True, -1)):
    f = te.make_xcch(n, 1000 + tsc)
    t = time.time(); want = o.xcch_send_frames(f, lsb, tsc); tr = time.time() - t
    got = dsp.xcch_encode_host(f, lsb, tsc)
    same = np.array_equal(got, want); ok &= same
    print("xcch  %7d frames  lsb8msb=%d tsc=%2d  %s  (reference %.1f s)" % (n, lsb, tsc, "identical" if same else "DIFFER", tr))
d, f, steal = te.make_tch(n, 2000, p_steal=0.2)
for lsb, tsc, splits in ((True, 3, [(0, n)]), (False, 6, [(0, 7), (7, 8), (8, 9), (9, 40), (40, n // 2), (n // 2, n)])):
    want = te.ref_tch(o, d, f, steal, lsb, tsc, splits)
    got, carry = te.our_tch(dsp.tch_encode_host, d, f, steal, lsb, tsc, splits)
    same = np.array_equal(got, want); ok &= same
    print("tch   %7d blocks  lsb8msb=%d tsc=%2d  %d call(s)  %s" % (n, lsb, tsc, len(splits), "identical" if same else "DIFFER"))
print("ALL IDENTICAL" if ok else "MISMATCH")
