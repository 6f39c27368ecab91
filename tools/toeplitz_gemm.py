#!/usr/bin/env python3
"""Measurement aid (NOT a product path): the RX 65/96 polyphase resampler written as a Toeplitz GEMM on the tensor cores,
timed against k_resample_rx_v3, with the damage it does to the demodulator's decisions counted.

north_star keeps "a tensor-core Toeplitz-GEMM formulation of the FIR only if ncu shows it beats the CUDA-core version".
Formulation: the stream is a matrix X of periods x 192 floats (96 interleaved complex samples per row, contiguous, no
copy); output period G (65 complex = 130 floats) = X[G-2] H0 + X[G-1] H1 + X[G] H2 + X[G+1] H3, each H_b a 192 x 130
matrix holding the 961 real taps at the places sigProcLib.cpp:1157-1210 reads them (re -> re, im -> im); the last period
of every 9-period chunk uses matrices with the taps the reference cannot see there removed (:1183-1186).  A dense GEMM
spends 2 x 4 x 192 x 130 = 199 680 flops per period where the direct form needs 65 x 14.8 x 4 = 3 844: the tensor cores
have to be 52x faster per flop just to break even, and the result is no longer the reference's sum order (nor, in TF32,
its precision).  Variants: cuBLAS TF32 (tensor cores), TF32 x3 split (hi/lo, near-FP32 products), cuBLAS FP32 (CUDA
cores, FP32 but a different summation order).  Each is timed with CUDA events and its resampled stream is pushed through
the product's detect + equalise kernels; mismatching flags / TOA / hard bits / soft bytes are counted against the exact path.

    python tools/toeplitz_gemm.py [blocks]     ->  one JSON object on stdout
"""
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
nper = nch * 9
dev = torch.device("cuda:0")
st = torch.cuda.current_stream()
dsp = pkg.BtsDsp(0, 1)


def h_blocks(taps, last_of_chunk):
    """4 matrices (192 x 130): block b multiplies the row of period G - 2 + b"""
    H = np.zeros((4, 192, 130), np.float32)
    for r in range(65):
        o = r + 135
        br, ix = (96 * o) % 65, (96 * o) // 65            # window index of the first tap, relative to sample 96 G - 192
        k = 0
        while br + 65 * k < 961:
            j = ix - k
            if j < 0:
                break
            if not (last_of_chunk and j >= 288):          # samples of the next chunk are invisible to the reference
                b, c = divmod(j, 96)
                H[b, 2 * c, 2 * r] = taps[br + 65 * k]
                H[b, 2 * c + 1, 2 * r + 1] = taps[br + 65 * k]
            k += 1
    return H


taps = dsp.table(9)
assert taps.size == 961
Hn = torch.from_numpy(h_blocks(taps, False)).to(dev)
Hl = torch.from_numpy(h_blocks(taps, True)).to(dev)

# ---- the bench's stream
g = torch.Generator(device=dev)
g.manual_seed(0xB2000002)
bits = torch.randint(0, 2, (nb, 148), generator=g, device=dev, dtype=torch.uint8)
bits[:, :3] = 0
bits[:, 145:] = 0
bits[:, 61:87] = torch.from_numpy(synth.bits_of(synth.TSC[0]).copy()).to(dev)
iq = torch.empty(nch * 864 * 2, dtype=torch.int16, device=dev)
dsp.tx_stream_dev(bits, nb, iq, stream=st)
raw = iq.to(torch.float32)
del iq
raw.add_(torch.randn(raw.numel(), generator=g, device=dev), alpha=955.0)
raw.round_().clamp_(-32768, 32767)

# X with two zero periods in front (the stream's zero history) and one behind
X = torch.zeros((nper + 3, 192), device=dev)
X[2:2 + nper] = raw.view(nper, 192)
out = torch.empty((nper, 130), device=dev)
outl = torch.empty((nch, 130), device=dev)


def gemm_pass(mm):
    """out[G] = sum_b X[G + b] Hn[b] (X is shifted by two rows), then the chunk-closing periods again with Hl"""
    torch.mm(X[0:nper], Hn[0], out=out)
    for b in (1, 2, 3):
        mm(out, X[b:b + nper], Hn[b])
    Xl = X[8:8 + nper]                                       # rows of the periods G = 9 c + 8: every 9th row
    torch.mm(Xl[0::9], Hl[0], out=outl)
    for b in (1, 2, 3):
        mm(outl, X[8 + b:8 + b + nper][0::9], Hl[b])
    out.view(nch, 9, 130)[:, 8] = outl


def plain(o, a, h):
    o.addmm_(a, h)


def split3(o, a, h):
    """TF32 x3: a = a_hi + a_lo, h = h_hi + h_lo with the hi parts exactly representable in TF32; three products"""
    a_hi = (a.view(torch.int32) & ~0x1FFF).view(torch.float32)
    h_hi = (h.view(torch.int32) & ~0x1FFF).view(torch.float32)
    o.addmm_(a_hi, h_hi)
    o.addmm_(a - a_hi, h_hi)
    o.addmm_(a_hi, h - h_hi)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps):
        fn()
    b.record(st)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def demod(res):
    tsc = torch.zeros(nb, dtype=torch.uint8, device=dev)
    flag = torch.zeros(nb, dtype=torch.int32, device=dev)
    amp = torch.zeros(nb * 2, device=dev)
    toa = torch.zeros(nb, device=dev)
    u8 = torch.zeros(nb * 148, dtype=torch.uint8, device=dev)
    dsp.demod_normal_u8_dev(res, 0, tsc, nb, flag, amp, toa, u8, 148, stream=st)
    torch.cuda.synchronize()
    return flag, toa, u8.view(nb, 148)


# ---- the exact path
res = torch.empty(nch * 585 * 2, device=dev)
ms_exact = timeit(lambda: dsp.resample_rx_dev(raw, nch, res, stream=st))
f0, t0, u0 = demod(res)
exact = res.view(nper, 130).clone()
result = {"blocks": blocks, "bursts": nb, "chunks": nch, "exact_k_resample_rx_v3_ms": ms_exact,
          "direct_flops_per_period": 65 * 961 * 4 // 65, "gemm_flops_per_period": 2 * 4 * 192 * 130, "variants": {}}

for name, tf32, mm in (("cublas_tf32", True, plain), ("cublas_tf32x3_split", True, split3), ("cublas_fp32", False, plain)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    ms = timeit(lambda: gemm_pass(mm))
    gemm_pass(mm)
    torch.cuda.synchronize()
    err = (out - exact).abs()
    f1, t1, u1 = demod(out.view(-1))
    both = (f0 != 0) & (f1 != 0)
    hard0, hard1 = u0 > 127, u1 > 127
    result["variants"][name] = {
        "ms": ms, "vs_exact_kernel": ms / ms_exact,
        "resampled_samples_differing": float((out != exact).float().mean()),
        "max_abs_error": float(err.max()), "rms_signal": float(exact.pow(2).mean().sqrt()),
        "flag_mismatches": int((f0 != f1).sum()), "toa_mismatches": int((t0 != t1).sum()),
        "bursts_with_a_different_hard_bit": int(((hard0 != hard1).any(dim=1) & both).sum()),
        "hard_bits_different": int((hard0 != hard1)[both].sum()),
        "bursts_with_a_different_soft_byte": int(((u0 != u1).any(dim=1) & both).sum()),
    }
torch.backends.cuda.matmul.allow_tf32 = False
print(json.dumps(result, indent=1))
