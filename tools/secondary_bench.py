#!/usr/bin/env python3
"""Throughput of the kernels outside the headline bench (configs 3-5): RACH detect+demod, pitched normal-burst
demod (1024 ARFCN x 8 TS per launch), GMSK modulate, TX resample.  Device-resident, CUDA events.  Measurement aid."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import openbts_ttsou_b200 as pkg
import synth

dev = torch.device("cuda:0")
dsp = pkg.BtsDsp(0, 1)
st = torch.cuda.current_stream()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps):
        fn()
    b.record(st)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {}
g = torch.Generator(device=dev); g.manual_seed(5)
# --- TX: 936*64 bursts -> modulate -> TX resample
nb = 936 * 256
bits = torch.randint(0, 2, (nb, 148), generator=g, device=dev, dtype=torch.uint8)
bits[:, 61:87] = torch.from_numpy(synth.bits_of(synth.TSC[0]).copy()).to(dev)
stream = torch.zeros(nb // 4 * 625 * 2, device=dev)
nch = nb // 4 * 625 // 585
iq = torch.zeros(nch * 864 * 2, dtype=torch.int16, device=dev)
ms = timeit(lambda: dsp.modulate_dev(bits, 148, nb, stream, 0, stream=st))
out["modulate"] = {"bursts": nb, "ms": ms, "bursts_per_s": nb / ms * 1e3, "gbs": nb * 1398 / ms / 1e6}
ms = timeit(lambda: dsp.resample_tx_dev(stream, nch, iq, stream=st))
out["resample_tx"] = {"chunks": nch, "ms": ms, "burst_eq_per_s": nb / ms * 1e3, "gbs": nch * 8136 / ms / 1e6}
ms = timeit(lambda: dsp.tx_stream_dev(bits, nb, iq, stream=st))
out["tx_fused_bits_to_int16"] = {"bursts": nb, "ms": ms, "bursts_per_s": nb / ms * 1e3, "gbs": (nb * 148 + nch * 3456) / ms / 1e6}
# --- RX at the wire formats, device-resident: int16 ingest resample and soft-byte demod over the same 64 000 chunks
iq_rx = (iq.to(torch.float32) + 400.0 * torch.randn(iq.numel(), generator=g, device=dev)).round().clamp_(-32768, 32767).to(torch.int16)
res16 = torch.zeros(nch * 585 * 2, device=dev)
ms = timeit(lambda: dsp.resample_rx_i16_dev(iq_rx, nch, res16, stream=st))
out["resample_rx_i16"] = {"chunks": nch, "ms": ms, "gbs": nch * 8136 / ms / 1e6}
resf = torch.zeros(nch * 585 * 2, device=dev)
rawf = iq_rx.to(torch.float32)
ms = timeit(lambda: dsp.resample_rx_dev(rawf, nch, resf, stream=st))
out["resample_rx_f32"] = {"chunks": nch, "ms": ms, "gbs": nch * 11592 / ms / 1e6}
nbw = nch * 585 // 625 * 4
tscw = torch.zeros(nbw, dtype=torch.uint8, device=dev)
fw = torch.zeros(nbw, dtype=torch.int32, device=dev); aw = torch.zeros(nbw * 2, device=dev); tw = torch.zeros(nbw, device=dev)
u8w = torch.zeros(nbw * 148, dtype=torch.uint8, device=dev)
ms = timeit(lambda: dsp.demod_normal_u8_dev(res16, 0, tscw, nbw, fw, aw, tw, u8w, 148, stream=st))
out["demod_normal_u8_stream"] = {"bursts": nbw, "ms": ms, "bursts_per_s": nbw / ms * 1e3, "detected": float(fw.float().mean())}
# --- config 4: 8192 pitched bursts per launch (mixed TSC), from a stream demod of the TX signal re-cut to pitch 160
raw = iq.to(torch.float32) + 400.0 * torch.randn(iq.numel(), generator=g, device=dev)
res = torch.zeros(nch * 585 * 2, device=dev)
dsp.resample_rx_dev(raw, nch, res, stream=st)
torch.cuda.synchronize()
n4 = 8192
host = res.cpu().numpy().view(np.complex64)
pitched = np.zeros((n4, 160), np.complex64)
for i in range(n4):
    o = (i // 4) * 625 + (0, 157, 313, 469)[i % 4]
    pitched[i, :157 if i % 4 == 0 else 156] = host[o:o + (157 if i % 4 == 0 else 156)]
dp = torch.from_numpy(pitched.view(np.float32).copy()).to(dev)
tsc = torch.zeros(n4, dtype=torch.uint8, device=dev)
flag = torch.zeros(n4, dtype=torch.int32, device=dev); amp = torch.zeros(n4 * 2, device=dev)
toa = torch.zeros(n4, device=dev); soft = torch.zeros(n4 * 148, device=dev)
ms = timeit(lambda: dsp.demod_normal_dev(dp, 160, tsc, n4, flag, amp, toa, soft, 148, stream=st), reps=50)
out["demod_normal_8192_per_launch"] = {"bursts": n4, "ms": ms, "bursts_per_s": n4 / ms * 1e3, "detected": float(flag.float().mean())}
# --- config 3: access bursts
nr = 524288
rb = torch.zeros(nr, 160, 2, device=dev)
rbits = np.zeros(88, np.uint8); rbits[:8] = synth.bits_of(synth.RACH_EXT_TAIL); rbits[8:49] = synth.bits_of(synth.RACH_SYNC)
x = dsp.modulate(rbits, 156 - 88)
xr = torch.from_numpy(np.stack([x.real, x.imag], -1).astype(np.float32)).to(dev) * 1000.0
rb[:, :156] = xr[None] + 100.0 * torch.randn(nr, 156, 2, generator=g, device=dev)
rflag = torch.zeros(nr, dtype=torch.int32, device=dev); ramp = torch.zeros(nr * 2, device=dev)
rtoa = torch.zeros(nr, device=dev); rsoft = torch.zeros(nr * 160, device=dev)
ms = timeit(lambda: dsp.rach_dev(rb, 160, nr, rflag, ramp, rtoa, rsoft, 160, stream=st), reps=5)
out["rach_detect_demod"] = {"bursts": nr, "ms": ms, "bursts_per_s": nr / ms * 1e3, "detected": float(rflag.float().mean()),
                            "gbs": nr * 1272 / ms / 1e6}
ms = timeit(lambda: dsp.rach_dev(rb, 160, nr, rflag, ramp, rtoa, None, 160, stream=st), reps=5)
out["rach_detect_only"] = {"bursts": nr, "ms": ms, "bursts_per_s": nr / ms * 1e3}
# --- caller policy (pullRadioVector semantics over batches): 1024 ARFCN x 8 TN x 32 frames per pull, all TSC slots
#     except TN0 = combination V (RACH on most frames); bursts = the pitched normal bursts above, tiled
A, F = 1024, 32
npol = A * 8 * F
pb = dp.view(n4, 320).repeat(npol // n4, 1).contiguous()
ct = np.ones((A, 8), np.uint8); ct[:, 0] = 5
trx = dsp.trx_create(np.zeros(A, np.uint8), ct, 0)
pvalid = torch.zeros(npol, dtype=torch.int32, device=dev)
pdg = torch.zeros(npol * 160, dtype=torch.uint8, device=dev)
fnc = [0]
def pull():
    dsp.trx_pull_dev(trx, pb, 160, F, fnc[0], pvalid, pdg, 160, stream=st)
    fnc[0] += F
ms = timeit(pull, reps=10)
out["trx_pull_policy"] = {"bursts": npol, "arfcn": A, "frames_per_pull": F, "ms": ms, "bursts_per_s": npol / ms * 1e3,
                          "valid": float(pvalid.float().mean())}
dsp.trx_destroy(trx)
# --- L1 FEC after the path: XCCH block decode, 65 536 frames (262 144 bursts) of soft bytes
nfr = 65536
sb = torch.randint(0, 256, (nfr * 4, 148), generator=g, device=dev, dtype=torch.uint8)
fu = torch.zeros(nfr * 228, dtype=torch.uint8, device=dev); fok = torch.zeros(nfr, dtype=torch.int32, device=dev)
ms = timeit(lambda: dsp.xcch_decode_dev(sb, 148, nfr, fu, fok, stream=st), reps=10)
out["xcch_decode"] = {"frames": nfr, "bursts": nfr * 4, "ms": ms, "bursts_per_s": nfr * 4 / ms * 1e3}
print(json.dumps(out, indent=1))
