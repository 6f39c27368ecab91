#!/usr/bin/env python3
"""One-off wide parity sweep on the GPU box: many random normal / access bursts and a long stream through the C ABI vs
the compiled reference, comparing BIT PATTERNS (so a -0 / +0 difference would show too).  Measurement aid."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import openbts_ttsou_b200 as pkg
from oracle.oracle import Oracle
import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # shifts both generators' seeds (default: the r2v/r4 sweep)
o = Oracle("best")
dsp = pkg.BtsDsp(0, 1)
def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8)
def cmp(name, got, want):
    same_val = np.array_equal(got, want)
    same_bits = bits(got).tobytes() == bits(want).tobytes()
    print("%-28s values %s  bit patterns %s" % (name, "identical" if same_val else "DIFFER", "identical" if same_bits else "differ (zero signs)" if same_val else "DIFFER"))
    return same_val
ok = True
t = time.time()
bursts, lens, tsc, _ = synth.make_normal_batch(o.modulate, n, seed=101 + seed, noise_only=0.1)
print("made %d normal bursts in %.1f s (seed shift %d)" % (n, time.time() - t, seed))
ref = o.rx_normal_batch(bursts, lens, tsc, threads=16)
got = dsp.demod_normal_host(bursts, lens, tsc, debug=True) if hasattr(dsp, "demod_normal_host") else None
for k in ("flag", "amp", "toa", "chan", "off", "w", "b", "soft"):
    g = got[k] if k != "soft" else got[k][:, :ref[k].shape[1]]
    w = ref[k] if k != "soft" else ref[k][:, :g.shape[1]]
    ok &= cmp("normal " + k, g, w)
rb, rl, _, _ = synth.make_rach_batch(o.modulate, n // 2, seed=202 + seed)
ref = o.rx_rach_batch(rb, rl, threads=16)
got = dsp.rach_host(rb, rl)
for k in ("flag", "amp", "toa", "soft"):
    g = got[k][:, :157] if k == "soft" else got[k]
    w = ref[k][:, :157] if k == "soft" else ref[k]
    ok &= cmp("rach " + k, g, w)
print("ALL IDENTICAL" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
