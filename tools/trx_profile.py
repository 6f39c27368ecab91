#!/usr/bin/env python3
"""Three policy pulls (1024 ARFCN x 8 TN x 32 frames) for an ncu launch list.  Measurement aid."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import openbts_ttsou_b200 as pkg
import synth
dev = torch.device("cuda:0")
dsp = pkg.BtsDsp(0, 1)
A, F = 1024, 32
n = A * 8 * F
g = torch.Generator(device=dev); g.manual_seed(1)
nb = 936 * 288
bits = torch.randint(0, 2, (nb, 148), generator=g, device=dev, dtype=torch.uint8)
bits[:, 61:87] = torch.from_numpy(synth.bits_of(synth.TSC[0]).copy()).to(dev)
nch = nb // 4 * 625 // 585
iq = torch.zeros(nch * 864 * 2, dtype=torch.int16, device=dev)
dsp.tx_stream_dev(bits, nb, iq)
raw = iq.to(torch.float32) + 400.0 * torch.randn(iq.numel(), generator=g, device=dev)
res = torch.zeros(nch * 585 * 2, device=dev)
dsp.resample_rx_dev(raw, nch, res)
ct = np.ones((A, 8), np.uint8); ct[:, 0] = 5
trx = dsp.trx_create(np.zeros(A, np.uint8), ct, 0)
valid = torch.zeros(n, dtype=torch.int32, device=dev)
dg = torch.zeros(n * 160, dtype=torch.uint8, device=dev)
for k in range(3):
    dsp.trx_pull_streams_dev(trx, res, F * 1250, F, k * F, valid, dg, 160)
torch.cuda.synchronize()
print("valid", float(valid.float().mean()))
