"""RSSI of the RX datagram, (int) floor(20.0*log10(9450.0/|amp|)) (reference Transceiver.cpp:400): the device must give
the integer the HOST's libm gives for every float amplitude, including the floats right at each floor boundary, where a
one-ulp difference between CUDA's log10 and glibc's would flip a datagram byte."""
import ctypes
import math

import numpy as np
import pytest

_libm = ctypes.CDLL("libm.so.6")
_libm.log10.restype = ctypes.c_double
_libm.log10.argtypes = [ctypes.c_double]


def glibc_rssi(a):
    """the reference's expression with glibc's log10 (what oracle/ref_shim.cpp's datagram glue evaluates)"""
    return int(math.floor(20.0 * _libm.log10(9450.0 / float(np.float32(a)))))


def boundary_floats(thr):
    """every table threshold and the 4 floats either side of it, plus a log-spaced sweep of ordinary amplitudes"""
    bits = thr.view(np.uint32).astype(np.int64)
    around = (bits[:, None] + np.arange(-4, 5)[None, :]).reshape(-1)
    around = around[(around > 0) & (around <= 0x7F7FFFFF)].astype(np.uint32).view(np.float32)
    rng = np.random.default_rng(3)
    sweep = np.exp(rng.uniform(np.log(1e-3), np.log(1e7), 20000)).astype(np.float32)
    return np.concatenate([around, sweep, np.float32([1.0, 9450.0, 9449.9995, 9450.001, 250.0, 32767.0])])


def test_rssi_table_reproduces_glibc_at_every_boundary(hostemu):
    thr = np.zeros(1544, np.float32)
    hostemu.emu_rssi.argtypes = [ctypes.c_float]
    rmin = hostemu.emu_rssi_table(thr.ctypes.data_as(ctypes.c_void_p))
    assert rmin == -700
    assert np.all(np.diff(thr.view(np.uint32).astype(np.int64)) <= 0), "thresholds must not increase with RSSI"
    vals = boundary_floats(thr[np.isfinite(thr) & (thr > 0)])
    bad = [(float(a), hostemu.emu_rssi(float(a)), glibc_rssi(a)) for a in vals if hostemu.emu_rssi(float(a)) != glibc_rssi(a)]
    assert not bad, bad[:5]
    # the amplitudes a receiver sees (RSSI -20 .. +60 dB) cross every integer in between
    seen = {hostemu.emu_rssi(float(a)) for a in vals}
    assert set(range(-20, 61)) <= seen


@pytest.mark.gpu
def test_rssi_on_the_gpu_equals_glibc(dsp, hostemu):
    thr = np.zeros(1544, np.float32)
    hostemu.emu_rssi_table(thr.ctypes.data_as(ctypes.c_void_p))
    vals = boundary_floats(thr[np.isfinite(thr) & (thr > 0)])
    got = dsp.trx_rssi(vals)
    want = np.array([glibc_rssi(a) for a in vals], np.int32)
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, (vals[bad[:5]], got[bad[:5]], want[bad[:5]])
