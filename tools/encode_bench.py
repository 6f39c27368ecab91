#!/usr/bin/env python3
"""Throughput of the transmit-side L1 encoders on the GPU box (device-resident, CUDA events): XCCH frames -> bursts and one
traffic channel's blocks -> bursts.  Measurement aid; prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import openbts_ttsou_b200 as pkg  # noqa: E402

dsp = pkg.BtsDsp(0, 1)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
n = 1 << 20
frames = torch.randint(0, 2, (n, 184), dtype=torch.uint8, device=dev, generator=g)
bursts = torch.empty((4 * n, 148), dtype=torch.uint8, device=dev)
d260 = torch.randint(0, 2, (n, 260), dtype=torch.uint8, device=dev, generator=g)
steal = (torch.rand(n, device=dev, generator=g) < 0.1).to(torch.uint8)
tb = torch.empty((4 * n + 4, 148), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream()


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


x = timeit(lambda: dsp.xcch_encode_dev(frames, n, 1, 2, bursts, 148, st))
t = timeit(lambda: dsp.tch_encode_dev(d260, frames, steal, n, 1, 5, None, tb, 148, st))
print(json.dumps({"xcch_encode": {"frames": n, "ms": x, "bursts_per_s": 4 * n / x * 1e3, "gbs": n * (184 + 592) / x / 1e6},
                  "tch_encode": {"blocks": n, "ms": t, "bursts_per_s": 4 * n / t * 1e3, "gbs": (float((steal == 0).sum()) * 260 + float((steal != 0).sum()) * 184 + n * (1 + 592)) / t / 1e6}}))
