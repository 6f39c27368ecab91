#!/usr/bin/env python3
"""Measurement aid: btsdsp_rx_stream_dev (resample -> detect -> equalise, device-resident) as one launch of each kernel
versus the segmented pipeline that runs the HBM-bound resampler of segment s+1 on a side stream, on a capped number of
SMs, next to the FP32-bound demod kernels of segment s.  Sweeps segment size x CTA cap, checks the outputs stay
bit-identical, prints one JSON line per point.   python tools/overlap_sweep.py [blocks]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import openbts_ttsou_b200 as pkg  # noqa: E402
import synth  # noqa: E402

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 855
nb, nch = blocks * 936, blocks * 250
dev = torch.device("cuda:0")
st = torch.cuda.current_stream()


def make(dsp):
    g = torch.Generator(device=dev)
    g.manual_seed(0xB2000002)
    bits = torch.randint(0, 2, (nb, 148), generator=g, device=dev, dtype=torch.uint8)
    bits[:, :3] = 0
    bits[:, 145:] = 0
    bits[:, 61:87] = torch.from_numpy(synth.bits_of(synth.TSC[0]).copy()).to(dev)
    iq = torch.empty(nch * 864 * 2, dtype=torch.int16, device=dev)
    dsp.tx_stream_dev(bits, nb, iq, stream=st)
    raw = iq.to(torch.float32)
    raw.add_(torch.randn(raw.numel(), generator=g, device=dev), alpha=955.0)
    return raw


def run(seg, ctas, raw=None, ref=None, light=1):
    os.environ["BTSDSP_RX_SEG"] = str(seg)
    os.environ["BTSDSP_RX_RES_CTAS"] = str(ctas)
    os.environ["BTSDSP_RX_LIGHT"] = str(light)
    dsp = pkg.BtsDsp(0, 1)
    if raw is None:
        raw = make(dsp)
    tsc = torch.zeros(nb, dtype=torch.uint8, device=dev)
    flag = torch.zeros(nb, dtype=torch.int32, device=dev)
    amp = torch.zeros(nb * 2, device=dev)
    toa = torch.zeros(nb, device=dev)
    soft = torch.zeros(nb * 148, device=dev)
    f = lambda: dsp.rx_stream_dev(raw, nch, tsc, nb, flag, amp, toa, soft, 148, stream=st)  # noqa: E731
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        a.record(st)
        for _ in range(10):
            f()
        b.record(st)
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 10)
    out = (flag.clone(), amp.clone(), toa.clone(), soft.clone())
    same = None if ref is None else all(torch.equal(x, y) for x, y in zip(out, ref))
    print(json.dumps({"seg_chunks": seg, "res_ctas": ctas, "light": light, "ms_per_step": round(best, 4), "bursts_per_s": nb / best * 1e3,
                      "identical_to_unsegmented": same}), flush=True)
    dsp.close()
    return raw, out


raw, ref = run(0, 0)
points = [(s, c) for s in (8000, 16000, 32000) for c in (0, 96, 64, 48, 32, 24)]
if len(sys.argv) > 2:
    points = [tuple(int(v) for v in p.split(":")) for p in sys.argv[2].split(",")]
for p in points:
    run(p[0], p[1], raw, ref, p[2] if len(p) > 2 else 1)
run(0, 0, raw, ref)
