"""Synthetic inputs of BASELINE.json's configs 3, 4 and 5, built ON THE DEVICE with the product's own
(parity-tested) modulator plus torch for the impairments, so that 10^6-burst batches take a second to make.  How an input
was made does not matter for parity: the same array is fed to the CUDA path and to the compiled reference.

Measurement infrastructure for bench.py / tools; not part of the product.
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402

PITCH = 160


def _delay_rows(x, d):
    """band-limited delay of every row of complex x (n, PITCH) by d[n] samples (circular over the row: the rows end in
    guard samples, so what wraps is pulse tail)"""
    n = x.shape[1]
    k = torch.fft.fftfreq(n, device=x.device) * n
    ph = torch.exp(-2j * math.pi * (k[None, :] * d[:, None]) / n)
    return torch.fft.ifft(torch.fft.fft(x, dim=1) * ph, dim=1)


def _lens(n, dev):
    return torch.where(torch.arange(n, device=dev) % 4 == 0, 157, 156)


def _finish(y, lens):
    """zero the samples past each slot's length, return float32 pairs (n, PITCH, 2)"""
    col = torch.arange(y.shape[1], device=y.device)[None, :]
    y = torch.where(col < lens[:, None], y, torch.zeros((), dtype=y.dtype, device=y.device))
    return torch.view_as_real(y.to(torch.complex64)).contiguous()


def rach_sweep(dsp, dev, per_cell=601, seed=0xB2000003, stream=None):
    """config 3: 64 TOA x 26 SNR x per_cell access bursts (per_cell 601 -> 1 000 064).  ext-tail + 41-bit sync + 36 random
    bits + 000, amp 1000 e^{j phi}, delay d + U[0,1), d = 0..63, SNR -5..20 dB, slot length 157/156/156/156.
    Returns (bursts (n,160,2) f32, toa_int (n,), snr_db (n,))."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n = 64 * 26 * per_cell
    idx = torch.arange(n, device=dev)
    d_int = (idx // per_cell) % 64                       # cell = (snr index, toa): bursts of one cell are adjacent
    snr = ((idx // per_cell) // 64).to(torch.float32) - 5.0
    perm = torch.randperm(n, generator=g, device=dev)    # ... then shuffled, so a launch sees every cell at once
    d_int, snr = d_int[perm], snr[perm]
    bits = torch.zeros((n, 88), dtype=torch.uint8, device=dev)
    bits[:, 0:8] = torch.from_numpy(synth.bits_of(synth.RACH_EXT_TAIL).copy()).to(dev)
    bits[:, 8:49] = torch.from_numpy(synth.bits_of(synth.RACH_SYNC).copy()).to(dev)
    bits[:, 49:85] = torch.randint(0, 2, (n, 36), generator=g, device=dev, dtype=torch.uint8)
    x = torch.zeros((n, PITCH, 2), dtype=torch.float32, device=dev)
    dsp.modulate_dev(bits, 88, n, x, PITCH, guard=68, first=0, stream=stream)
    torch.cuda.synchronize()
    lens = _lens(n, dev)
    out = torch.empty_like(x)
    step = 131072
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        m = hi - lo
        xc = torch.view_as_complex(x[lo:hi])
        d = d_int[lo:hi].to(torch.float32) + torch.rand(m, generator=g, device=dev)
        y = _delay_rows(xc, d)
        phi = 2 * math.pi * torch.rand(m, generator=g, device=dev)
        y = y * (1000.0 * torch.exp(1j * phi))[:, None]
        sigma = 1000.0 / torch.sqrt(10.0 ** (snr[lo:hi] / 10.0))
        noise = torch.view_as_complex(torch.randn((m, PITCH, 2), generator=g, device=dev)) * (sigma / math.sqrt(2.0))[:, None]
        out[lo:hi] = _finish(y + noise, lens[lo:hi])
    return out, d_int, snr


def normal_batch(dsp, dev, n_arfcn=1024, frames=128, seed=0xB2000004, empty=0.05, stream=None):
    """config 4: frames x n_arfcn x 8 TN normal bursts, row = (frame*n_arfcn + arfcn)*8 + tn; TSC = arfcn mod 8, amp
    log-uniform [500, 8000] x random phase, delay U[0,3) symbols, half of the ARFCNs behind a 2-tap channel
    [1, 0.4 e^{j theta}], SNR U[10,30] dB, `empty` of the slots noise only.
    Returns (bursts (n,160,2) f32, tsc (n,) u8, bits (n,148) u8, occupied (n,) bool)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n = frames * n_arfcn * 8
    row = torch.arange(n, device=dev)
    arfcn = (row // 8) % n_arfcn
    tsc = (arfcn % 8).to(torch.uint8)
    bits = torch.randint(0, 2, (n, 148), generator=g, device=dev, dtype=torch.uint8)
    bits[:, :3] = 0
    bits[:, 145:] = 0
    tsc_bits = torch.from_numpy(np.stack([synth.bits_of(t) for t in synth.TSC]).copy()).to(dev)
    bits[:, 61:87] = tsc_bits[tsc.long()]
    x = torch.zeros((n, PITCH, 2), dtype=torch.float32, device=dev)
    dsp.modulate_dev(bits, 148, n, x, PITCH, guard=-1, first=0, stream=stream)
    torch.cuda.synchronize()
    lens = _lens(n, dev)
    gch = torch.Generator(device=dev)
    gch.manual_seed(seed + 1)
    has_ch = torch.rand(n_arfcn, generator=gch, device=dev) < 0.5
    theta = 2 * math.pi * torch.rand(n_arfcn, generator=gch, device=dev)
    c2 = torch.where(has_ch, 0.4 * torch.exp(1j * theta), torch.zeros((), dtype=torch.complex64, device=dev))
    out = torch.empty_like(x)
    occupied = torch.rand(n, generator=g, device=dev) >= empty
    step = 131072
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        m = hi - lo
        xc = torch.view_as_complex(x[lo:hi])
        y = xc + c2[arfcn[lo:hi]][:, None] * torch.cat([torch.zeros((m, 1), dtype=xc.dtype, device=dev), xc[:, :-1]], dim=1)
        y = _delay_rows(y, 3.0 * torch.rand(m, generator=g, device=dev))
        a = torch.exp(math.log(500.0) + (math.log(8000.0) - math.log(500.0)) * torch.rand(m, generator=g, device=dev))
        phi = 2 * math.pi * torch.rand(m, generator=g, device=dev)
        y = y * (a * torch.exp(1j * phi))[:, None]
        y = torch.where(occupied[lo:hi, None], y, torch.zeros((), dtype=y.dtype, device=dev))
        snr = 10.0 + 20.0 * torch.rand(m, generator=g, device=dev)
        sigma = a / torch.sqrt(10.0 ** (snr / 10.0))
        noise = torch.view_as_complex(torch.randn((m, PITCH, 2), generator=g, device=dev)) * (sigma / math.sqrt(2.0))[:, None]
        out[lo:hi] = _finish(y + noise, lens[lo:hi])
    return out, tsc, bits, occupied


FRAMES_PER_BLOCK, BURSTS_PER_ARFCN, CHUNKS_PER_ARFCN = 117, 936, 250    # one 117-frame block per ARFCN radio


def arfcn_bits(arfcns, dev, seed=0xB2000005):
    """config 5: the 936 bursts (117 frames x 8 TN) an ARFCN transmits, TSC = arfcn mod 8; the same for a given ARFCN on
    whatever rank makes them.  Returns (bits (len(arfcns), 936, 148) u8, tsc (len(arfcns)*936,) u8)."""
    tsc_bits = torch.from_numpy(np.stack([synth.bits_of(t) for t in synth.TSC]).copy()).to(dev)
    out = torch.empty((len(arfcns), BURSTS_PER_ARFCN, 148), dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev)
    for i, a in enumerate(arfcns):
        g.manual_seed(seed + 7919 * int(a))
        b = torch.randint(0, 2, (BURSTS_PER_ARFCN, 148), generator=g, device=dev, dtype=torch.uint8)
        b[:, :3] = 0
        b[:, 145:] = 0
        b[:, 61:87] = tsc_bits[int(a) % 8]
        out[i] = b
    tsc = torch.tensor([int(a) % 8 for a in arfcns], dtype=torch.uint8, device=dev).repeat_interleave(BURSTS_PER_ARFCN)
    return out, tsc


def arfcn_air(iq_tx, arfcns, dev, sigma=955.0, seed=0xB2000006):
    """the air interface between the TX and RX radios of every ARFCN: AWGN on the int16 radio samples (SNR 20 dB at the
    13500 TX scale), re-quantised to int16 as an ADC delivers them.  iq_tx: (len(arfcns), 250*864*2) int16."""
    out = torch.empty_like(iq_tx)
    g = torch.Generator(device=dev)
    for i, a in enumerate(arfcns):
        g.manual_seed(seed + 104729 * int(a))
        noise = torch.randn(iq_tx.shape[1], generator=g, device=dev) * sigma
        out[i] = (iq_tx[i].to(torch.float32) + noise).round_().clamp_(-32768, 32767).to(torch.int16)
    return out
