"""ctypes binding of libbtsdsp.so (include/btsdsp.h) -- the Python face used by tests/ and bench.py.

The product is the shared library; this module only marshals pointers.  Layer-1 methods mirror the
reference's sigProcLib.h functions on numpy arrays (host pointers); the *_dev methods take anything with
a `.data_ptr()` (torch CUDA tensors) or a raw integer device address, plus a CUDA stream handle; the
*_host methods run the batched pipelines over host buffers (numpy, or pinned torch tensors).
There is no CPU fallback: if the library is missing or no B200 is visible, construction raises.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbtsdsp.so")

FULL_SPAN, OVERLAP_ONLY, START_ONLY, WITH_TAIL, NO_DELAY = 0, 1, 2, 3, 4
T_COS, T_SIN, T_ROT, T_REVROT, T_PULSE, T_MID_SEQ, T_MID_META, T_RACH_SEQ, T_RACH_META, T_LPF_RX, T_LPF_TX = range(11)

_vp, _i, _ll, _f, _u = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_uint


class _cf32(ctypes.Structure):
    _fields_ = [("re", _f), ("im", _f)]


_SIGS = {
    "btsdsp_create": (_i, [ctypes.POINTER(_vp), _i, _i]),
    "btsdsp_destroy": (_i, [_vp]),
    "btsdsp_last_error": (ctypes.c_char_p, [_vp]),
    "btsdsp_version": (_i, []),
    "btsdsp_device": (_i, [_vp]),
    "btsdsp_sps": (_i, [_vp]),
    "btsdsp_get_table": (_i, [_vp, _i, _i, _vp, _i]),
    "btsdsp_launch_count": (_ll, [_vp]),
    "btsdsp_synchronize": (_i, [_vp]),
    "btsdsp_set_timing": (_i, [_vp, _i]),
    "btsdsp_get_timing": (_i, [_vp, _vp, _vp]),
    "btsdsp_set_copy_only": (_i, [_vp, _i]),
    "btsdsp_convolve": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _i, _i]),
    "btsdsp_correlate": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _i, _i]),
    "btsdsp_scale_vector": (_i, [_vp, _vp, _i, _i, _cf32]),
    "btsdsp_delay_vector": (_i, [_vp, _vp, _i, _f]),
    "btsdsp_peak_detect": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "btsdsp_interpolate_point": (_i, [_vp, _vp, _i, _f, _vp]),
    "btsdsp_energy_detect": (_i, [_vp, _vp, _i, _u, _f, _vp, _vp]),
    "btsdsp_modulate_burst": (_i, [_vp, _vp, _i, _i, _vp, _i]),
    "btsdsp_analyze_traffic_burst": (_i, [_vp, _vp, _i, _u, _f, _vp, _vp, _i, _vp, _vp, _vp]),
    "btsdsp_detect_rach_burst": (_i, [_vp, _vp, _i, _f, _vp, _vp, _vp]),
    "btsdsp_design_dfe": (_i, [_vp, _vp, _i, _f, _i, _vp, _vp]),
    "btsdsp_equalize_burst": (_i, [_vp, _vp, _i, _f, _vp, _i, _vp, _i, _vp]),
    "btsdsp_demodulate_burst": (_i, [_vp, _vp, _i, _cf32, _f, _vp]),
    "btsdsp_polyphase_resample": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i]),
    "btsdsp_polyphase_resample_taps": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _vp, _i]),
    "btsdsp_create_lpf": (_i, [_vp, _i, _f, _vp, _i]),
    "btsdsp_vector_op": (_i, [_vp, _i, _vp, _i, _i, _vp, _i, _cf32, _vp]),
    "btsdsp_modulate_dev": (_i, [_vp, _vp, _i, _ll, _i, _ll, _vp, _ll, _vp]),
    "btsdsp_resample_rx_dev": (_i, [_vp, _vp, _i, _ll, _vp, _vp]),
    "btsdsp_resample_tx_dev": (_i, [_vp, _vp, _i, _ll, _vp, _vp]),
    "btsdsp_resample_rx_i16_dev": (_i, [_vp, _vp, _i, _i, _ll, _vp, _vp]),
    "btsdsp_resample_rx_i16_streams_dev": (_i, [_vp, _vp, _ll, _i, _i, _i, _ll, _vp, _ll, _vp]),
    "btsdsp_demod_normal_u8_dev": (_i, [_vp, _vp, _ll, _vp, _ll, _vp, _ll, _f, _f, _f, _vp, _vp, _vp, _vp, _i, _vp]),
    "btsdsp_rx_stream_wire_host": (_i, [_vp, _vp, _i, _ll, _vp, _ll, _f, _f, _f, _vp, _vp, _vp, _vp]),
    "btsdsp_demod_normal_dev": (_i, [_vp, _vp, _ll, _vp, _ll, _vp, _ll, _f, _f, _f, _vp, _vp, _vp, _vp, _i, _vp, _vp,
                                     _vp, _vp, _vp]),
    "btsdsp_analyze_dev": (_i, [_vp, _vp, _ll, _vp, _ll, _vp, _ll, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "btsdsp_rach_dev": (_i, [_vp, _vp, _ll, _vp, _ll, _ll, _f, _vp, _vp, _vp, _vp, _i, _vp]),
    "btsdsp_design_dfe_dev": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, _vp]),
    "btsdsp_equalize_dev": (_i, [_vp, _vp, _ll, _vp, _ll, _ll, _vp, _vp, _vp, _vp, _i, _vp, _ll, _vp]),
    "btsdsp_demodulate_dev": (_i, [_vp, _vp, _ll, _vp, _ll, _ll, _vp, _vp, _vp, _i, _vp]),
    "btsdsp_rx_stream_host": (_i, [_vp, _vp, _ll, _vp, _ll, _f, _f, _f, _vp, _vp, _vp, _vp, _i]),
    "btsdsp_rx_stream_dev": (_i, [_vp, _vp, _ll, _vp, _ll, _f, _f, _f, _vp, _vp, _vp, _vp, _i, _vp]),
    "btsdsp_rx_stream_cont_dev": (_i, [_vp, _vp, _i, _ll, _vp, _ll, _f, _f, _f, _vp, _vp, _vp, _vp, _i, _vp]),
    "btsdsp_tx_stream_host": (_i, [_vp, _vp, _ll, _vp]),
    "btsdsp_tx_stream_dev": (_i, [_vp, _vp, _ll, _vp, _vp]),
    "btsdsp_tx_streams_dev": (_i, [_vp, _vp, _ll, _i, _vp, _vp]),
    "btsdsp_demod_normal_host": (_i, [_vp, _vp, _ll, _vp, _vp, _ll, _f, _f, _f, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp,
                                      _vp]),
    "btsdsp_rach_host": (_i, [_vp, _vp, _ll, _vp, _ll, _f, _vp, _vp, _vp, _vp, _i]),
    "btsdsp_trx_rssi": (_i, [_vp, _vp, _i, _vp]),
    "btsdsp_trx_create": (_i, [_vp, _i, _vp, _vp, _i, _vp]),
    "btsdsp_trx_destroy": (_i, [_vp, _vp]),
    "btsdsp_trx_set_slot": (_i, [_vp, _vp, _i, _i, _i]),
    "btsdsp_trx_state_bytes": (_i, []),
    "btsdsp_trx_get_state": (_i, [_vp, _vp, _i, _vp, _i]),
    "btsdsp_trx_pull_dev": (_i, [_vp, _vp, _vp, _ll, _i, _i, _vp, _vp, _i, _vp]),
    "btsdsp_trx_pull_host": (_i, [_vp, _vp, _vp, _ll, _i, _i, _vp, _vp, _i]),
    "btsdsp_trx_pull_streams_dev": (_i, [_vp, _vp, _vp, _ll, _i, _i, _vp, _vp, _i, _vp]),
    "btsdsp_trx_radio_host": (_i, [_vp, _vp, _vp, _ll, _ll, _i, _i, _vp, _vp, _i]),
    "btsdsp_graph_begin": (_i, [_vp, _vp, _vp]),
    "btsdsp_graph_end": (_i, [_vp, _vp, _vp]),
    "btsdsp_graph_launch": (_i, [_vp, _vp, _vp]),
    "btsdsp_graph_destroy": (_i, [_vp, _vp]),
    "btsdsp_tch_decode_dev": (_i, [_vp, _vp, _i, _ll, _vp, _vp, _vp, _vp, _vp, _vp]),
    "btsdsp_tch_decode_host": (_i, [_vp, _vp, _i, _ll, _vp, _vp, _vp, _vp, _vp]),
    "btsdsp_tx_datagrams_host": (_i, [_vp, _vp, _ll, _i, _i, _i, _vp, _vp, _vp]),
    "btsdsp_tx_datagrams_52m_host": (_i, [_vp, _vp, _ll, _i, _i, _i, _vp, _vp, _vp]),
    "btsdsp_trx_set_variant_52m": (_i, [_vp, _vp, _i, _i]),
    "btsdsp_xcch_decode_dev": (_i, [_vp, _vp, _i, _ll, _vp, _vp, _vp]),
    "btsdsp_xcch_encode_dev": (_i, [_vp, _vp, _ll, _i, _i, _vp, _i, _vp]),
    "btsdsp_xcch_encode_host": (_i, [_vp, _vp, _ll, _i, _i, _vp, _i]),
    "btsdsp_tch_encode_dev": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _vp, _vp, _i, _vp]),
    "btsdsp_tch_encode_host": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _vp, _vp, _i]),
    "btsdsp_xcch_decode_host": (_i, [_vp, _vp, _i, _ll, _vp, _vp]),
    "btsdsp_rach_decode_dev": (_i, [_vp, _vp, _i, _ll, _vp, _vp, _vp]),
    "btsdsp_rach_decode_host": (_i, [_vp, _vp, _i, _ll, _vp, _vp]),
    "btsdsp_analyze_52m_dev": (_i, [_vp, _vp, _ll, _vp, _ll, _vp, _ll, _f, _u, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "btsdsp_analyze_traffic_burst_52m": (_i, [_vp, _vp, _i, _u, _f, _u, _i, _vp, _vp, _vp, _vp, _vp]),
    "btsdsp_energy_detect_52m": (_i, [_vp, _vp, _i, _u, _f, _vp, _vp]),
    "btsdsp_host_alloc": (_vp, [ctypes.c_size_t]),
    "btsdsp_host_free": (None, [_vp]),
}

EXPORTS = sorted(_SIGS)
_lib = None


def load_library():
    """dlopen libbtsdsp.so and declare every prototype.  Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libbtsdsp.so is not built (run `python -m openbts_ttsou_b200.build`); "
                               "there is no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class BtsDspError(RuntimeError):
    pass


def _p(a):
    """address of a numpy array / torch tensor / int / None"""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.complex64)


def _stream(stream):
    if stream is None:
        return None
    return getattr(stream, "cuda_stream", stream)


class BtsDsp:
    """One context = one GPU + one samples-per-symbol setting (btsdsp_create)."""

    def __init__(self, device=0, sps=1):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.btsdsp_create(ctypes.byref(h), device, sps)
        if rc != 0:
            raise BtsDspError("btsdsp_create failed (%d): %s" % (rc, self.lib.btsdsp_last_error(None).decode()))
        self.h = h
        self.sps = sps
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.btsdsp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise BtsDspError("btsdsp error %d: %s" % (rc, self.lib.btsdsp_last_error(self.h).decode()))
        return rc

    @property
    def launch_count(self):
        return int(self.lib.btsdsp_launch_count(self.h))

    def synchronize(self):
        self._ck(self.lib.btsdsp_synchronize(self.h))

    def set_timing(self, enable=True):
        self._ck(self.lib.btsdsp_set_timing(self.h, int(enable)))

    def trx_rssi(self, abs_amp):
        a = np.ascontiguousarray(abs_amp, np.float32)
        out = np.zeros(a.size, np.int32)
        self._ck(self.lib.btsdsp_trx_rssi(self.h, _p(a), a.size, _p(out)))
        return out

    def set_copy_only(self, enable=True):
        """measurement aid: the host pipelines move their bytes but launch no kernels (copy roofline)"""
        self._ck(self.lib.btsdsp_set_copy_only(self.h, int(enable)))

    def get_timing(self):
        """(detect_ms, equalize_ms) of the last timed demod_normal_dev call"""
        t = np.zeros(2, np.float32)
        self._ck(self.lib.btsdsp_get_timing(self.h, _p(t[0:]), _p(t[1:])))
        return float(t[0]), float(t[1])

    def table(self, tid, idx=0):
        buf = np.zeros(4096, np.float32)
        n = self._ck(self.lib.btsdsp_get_table(self.h, tid, idx, _p(buf), buf.size))
        out = buf[:n].copy()
        if tid in (T_ROT, T_REVROT, T_PULSE, T_MID_SEQ, T_RACH_SEQ):
            return out.view(np.complex64)
        return out

    # ---- layer 1: the sigProcLib.h functions on host vectors -------------------------------------
    def _conv(self, fn, a, b, span, a_real, b_real):
        a, b = _c64(a), _c64(b)
        out = np.zeros(a.size + b.size + 2, np.complex64)
        n = self._ck(fn(self.h, _p(a), a.size, int(a_real), _p(b), b.size, int(b_real), _p(out), out.size, span))
        return out[:n].copy()

    def convolve(self, a, b, span, a_real=False, b_real=False):
        return self._conv(self.lib.btsdsp_convolve, a, b, span, a_real, b_real)

    def correlate(self, a, b, span, a_real=False, b_real=False):
        return self._conv(self.lib.btsdsp_correlate, a, b, span, a_real, b_real)

    def scale_vector(self, v, scale, real_only=False):
        v = _c64(v).copy()
        s = complex(scale)
        self._ck(self.lib.btsdsp_scale_vector(self.h, _p(v), v.size, int(real_only), _cf32(s.real, s.imag)))
        return v

    def delay_vector(self, v, delay):
        v = _c64(v).copy()
        self._ck(self.lib.btsdsp_delay_vector(self.h, _p(v), v.size, float(delay)))
        return v

    def peak_detect(self, v):
        v = _c64(v)
        pk = np.zeros(1, np.complex64)
        r = np.zeros(2, np.float32)
        self._ck(self.lib.btsdsp_peak_detect(self.h, _p(v), v.size, _p(pk), _p(r[0:]), _p(r[1:])))
        return pk[0], float(r[0]), float(r[1])

    def interpolate_point(self, v, ix):
        v = _c64(v)
        pk = np.zeros(1, np.complex64)
        self._ck(self.lib.btsdsp_interpolate_point(self.h, _p(v), v.size, float(ix), _p(pk)))
        return pk[0]

    def energy_detect(self, v, win, thr):
        v = _c64(v)
        avg = np.zeros(1, np.float32)
        det = np.zeros(1, np.int32)
        self._ck(self.lib.btsdsp_energy_detect(self.h, _p(v), v.size, int(win), float(thr), _p(avg), _p(det)))
        return bool(det[0]), float(avg[0])

    def modulate(self, bits, guard):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        out = np.zeros(self.sps * (bits.size + guard), np.complex64)
        n = self._ck(self.lib.btsdsp_modulate_burst(self.h, _p(bits), bits.size, guard, _p(out), out.size))
        assert n == out.size
        return out

    def analyze(self, burst, tsc, thr, request=True):
        burst = _c64(burst)
        amp = np.zeros(1, np.complex64)
        chan = np.zeros(6 * self.sps, np.complex64)
        r = np.zeros(2, np.float32)
        det = np.zeros(1, np.int32)
        self._ck(self.lib.btsdsp_analyze_traffic_burst(self.h, _p(burst), burst.size, int(tsc), float(thr), _p(amp),
                                                       _p(r[0:]), int(request), _p(chan), _p(r[1:]), _p(det)))
        return bool(det[0]), amp[0], float(r[0]), chan, float(r[1])

    def detect_rach(self, burst, thr):
        burst = _c64(burst)
        amp = np.zeros(1, np.complex64)
        toa = np.zeros(1, np.float32)
        det = np.zeros(1, np.int32)
        self._ck(self.lib.btsdsp_detect_rach_burst(self.h, _p(burst), burst.size, float(thr), _p(amp), _p(toa), _p(det)))
        return bool(det[0]), amp[0], float(toa[0])

    def design_dfe(self, chan, snr, nf=7):
        chan = _c64(chan)
        w = np.zeros(nf, np.complex64)
        b = np.zeros(max(chan.size - 1, 1), np.complex64)
        self._ck(self.lib.btsdsp_design_dfe(self.h, _p(chan), chan.size, float(snr), nf, _p(w), _p(b)))
        return w, b[:chan.size - 1]

    def equalize(self, burst, toa, w, b):
        """returns (soft, burst_after): like the reference, the burst is delayed in place"""
        burst = _c64(burst).copy()
        w, b = _c64(w), _c64(b)
        soft = np.zeros(burst.size, np.float32)
        self._ck(self.lib.btsdsp_equalize_burst(self.h, _p(burst), burst.size, float(toa), _p(w), w.size, _p(b), b.size,
                                                _p(soft)))
        return soft, burst

    def demodulate(self, burst, amp, toa):
        burst = _c64(burst)
        soft = np.zeros(burst.size, np.float32)
        a = complex(amp)
        n = self._ck(self.lib.btsdsp_demodulate_burst(self.h, _p(burst), burst.size, _cf32(a.real, a.imag), float(toa),
                                                      _p(soft)))
        return soft[:n].copy()

    def resample(self, x, P, Q, lpf):
        x = _c64(x)
        out = np.zeros(int(np.ceil(x.size * P / Q)) + 4, np.complex64)
        n = self._ck(self.lib.btsdsp_polyphase_resample(self.h, _p(x), x.size, P, Q, lpf, _p(out), out.size))
        return out[:n].copy()

    def resample_taps(self, x, P, Q, taps, real_taps=True):
        """polyphaseResampleVector with the caller's own filter (complex taps; real_taps: only their real parts count)"""
        x, taps = _c64(x), _c64(taps)
        out = np.zeros(int(np.ceil(x.size * P / Q)) + 4, np.complex64)
        n = self._ck(self.lib.btsdsp_polyphase_resample_taps(self.h, _p(x), x.size, P, Q, _p(taps), taps.size, int(real_taps),
                                                             _p(out), out.size))
        return out[:n].copy()

    def create_lpf(self, filter_len, gain_dc):
        out = np.zeros(961, np.float32)
        n = self._ck(self.lib.btsdsp_create_lpf(self.h, filter_len, gain_dc, _p(out), out.size))
        return out[:n].copy()

    def vector_op(self, op, x, real_only=False, y=None, scalar=0j):
        """op: 0 add, 1 offset, 2 conj, 3 slice, 5 rotate, 6 reverse-rotate -> the new x; 4 norm2 -> float"""
        x = _c64(x).copy()
        yy = _c64(y) if y is not None else None
        res = np.zeros(1, np.float32)
        self._ck(self.lib.btsdsp_vector_op(self.h, op, _p(x), x.size, int(real_only), _p(yy) if yy is not None else None,
                                           yy.size if yy is not None else 0, _cf32(scalar.real, scalar.imag), _p(res)))
        return float(res[0]) if op == 4 else x

    # ---- layer 2: device pointers ------------------------------------------------------------------
    def modulate_dev(self, bits, nbits, n, out, pitch, guard=-1, first=0, stream=None):
        self._ck(self.lib.btsdsp_modulate_dev(self.h, _p(bits), nbits, n, guard, first, _p(out), pitch, _stream(stream)))

    def resample_rx_dev(self, raw, nchunks, out, has_history=False, stream=None):
        self._ck(self.lib.btsdsp_resample_rx_dev(self.h, _p(raw), int(has_history), nchunks, _p(out), _stream(stream)))

    def resample_rx_i16_dev(self, iq, nchunks, out, swap_iq=False, has_history=False, stream=None):
        self._ck(self.lib.btsdsp_resample_rx_i16_dev(self.h, _p(iq), int(swap_iq), int(has_history), nchunks, _p(out),
                                                     _stream(stream)))

    def resample_rx_i16_streams_dev(self, iq, iq_pitch, nstreams, nchunks, out, out_pitch, swap_iq=False, has_history=False,
                                    stream=None):
        self._ck(self.lib.btsdsp_resample_rx_i16_streams_dev(self.h, _p(iq), iq_pitch, nstreams, int(swap_iq), int(has_history),
                                                             nchunks, _p(out), out_pitch, _stream(stream)))

    def demod_normal_u8_dev(self, bursts, pitch, tsc, n, flag, amp, toa, soft_u8, soft_pitch=148, lens=None, first=0,
                            detect_thr=3.0, gate_thr=-1.0, snr_thr=250.0, stream=None):
        self._ck(self.lib.btsdsp_demod_normal_u8_dev(self.h, _p(bursts), pitch, _p(lens), first, _p(tsc), n, detect_thr,
                                                     gate_thr, snr_thr, _p(flag), _p(amp), _p(toa), _p(soft_u8), soft_pitch,
                                                     _stream(stream)))

    def rx_stream_wire_host(self, iq, nchunks, tsc, nbursts, flag, amp, toa, soft_u8, swap_iq=False, detect_thr=3.0,
                            gate_thr=-1.0, snr_thr=250.0):
        self._ck(self.lib.btsdsp_rx_stream_wire_host(self.h, _p(iq), int(swap_iq), nchunks, _p(tsc), nbursts, detect_thr,
                                                     gate_thr, snr_thr, _p(flag), _p(amp), _p(toa), _p(soft_u8)))

    def resample_tx_dev(self, x, nchunks, out, has_history=False, stream=None):
        self._ck(self.lib.btsdsp_resample_tx_dev(self.h, _p(x), int(has_history), nchunks, _p(out), _stream(stream)))

    def demod_normal_dev(self, bursts, pitch, tsc, n, flag, amp, toa, soft, soft_pitch=160, lens=None, first=0,
                         detect_thr=3.0, gate_thr=-1.0, snr_thr=250.0, chan=None, off=None, w=None, b=None, stream=None):
        self._ck(self.lib.btsdsp_demod_normal_dev(self.h, _p(bursts), pitch, _p(lens), first, _p(tsc), n, detect_thr,
                                                  gate_thr, snr_thr, _p(flag), _p(amp), _p(toa), _p(soft), soft_pitch,
                                                  _p(chan), _p(off), _p(w), _p(b), _stream(stream)))

    def analyze_dev(self, bursts, pitch, tsc, n, flag, amp, toa, chan=None, off=None, lens=None, first=0,
                    detect_thr=3.0, request=True, stream=None):
        self._ck(self.lib.btsdsp_analyze_dev(self.h, _p(bursts), pitch, _p(lens), first, _p(tsc), n, detect_thr,
                                             int(request), _p(flag), _p(amp), _p(toa), _p(chan), _p(off), _stream(stream)))

    def rach_dev(self, bursts, pitch, n, flag, amp, toa, soft=None, soft_pitch=160, lens=None, first=0, detect_thr=5.0,
                 stream=None):
        self._ck(self.lib.btsdsp_rach_dev(self.h, _p(bursts), pitch, _p(lens), first, n, detect_thr, _p(flag), _p(amp),
                                          _p(toa), _p(soft), soft_pitch, _stream(stream)))

    def design_dfe_dev(self, chan, snr, n, w, b, stream=None):
        self._ck(self.lib.btsdsp_design_dfe_dev(self.h, _p(chan), _p(snr), n, _p(w), _p(b), _stream(stream)))

    def equalize_dev(self, bursts, pitch, n, toa, w, b, soft, soft_pitch=160, burst_out=None, out_pitch=0, lens=None,
                     first=0, stream=None):
        self._ck(self.lib.btsdsp_equalize_dev(self.h, _p(bursts), pitch, _p(lens), first, n, _p(toa), _p(w), _p(b),
                                              _p(soft), soft_pitch, _p(burst_out), out_pitch, _stream(stream)))

    def demodulate_dev(self, bursts, pitch, n, amp, toa, soft, soft_pitch=160, lens=None, first=0, stream=None):
        self._ck(self.lib.btsdsp_demodulate_dev(self.h, _p(bursts), pitch, _p(lens), first, n, _p(amp), _p(toa), _p(soft),
                                                soft_pitch, _stream(stream)))

    def rx_stream_dev(self, raw, nchunks, tsc, nbursts, flag, amp, toa, soft, soft_pitch=148, detect_thr=3.0,
                      gate_thr=-1.0, snr_thr=250.0, stream=None):
        self._ck(self.lib.btsdsp_rx_stream_dev(self.h, _p(raw), nchunks, _p(tsc), nbursts, detect_thr, gate_thr, snr_thr,
                                               _p(flag), _p(amp), _p(toa), _p(soft), soft_pitch, _stream(stream)))

    def rx_stream_cont_dev(self, raw, has_history, nchunks, tsc, nbursts, flag, amp, toa, soft, soft_pitch=148, detect_thr=3.0,
                           gate_thr=-1.0, snr_thr=250.0, stream=None):
        """a piece of a running stream (raw may be an address: the 192 samples before it must be valid when has_history)"""
        self._ck(self.lib.btsdsp_rx_stream_cont_dev(self.h, _p(raw), int(has_history), nchunks, _p(tsc), nbursts, detect_thr, gate_thr,
                                                    snr_thr, _p(flag), _p(amp), _p(toa), _p(soft), soft_pitch, _stream(stream)))

    def tx_stream_dev(self, bits148, n, out, stream=None):
        self._ck(self.lib.btsdsp_tx_stream_dev(self.h, _p(bits148), n, _p(out), _stream(stream)))

    def tx_streams_dev(self, bits148, n, nstreams, out, stream=None):
        """nstreams independent TX chains of n bursts each in one launch (bits and out laid out stream after stream)"""
        self._ck(self.lib.btsdsp_tx_streams_dev(self.h, _p(bits148), n, nstreams, _p(out), _stream(stream)))

    # ---- layer 3: host buffers ---------------------------------------------------------------------
    def rx_stream_host(self, raw, nchunks, tsc, nbursts, flag, amp, toa, soft, soft_pitch=148, detect_thr=3.0,
                       gate_thr=-1.0, snr_thr=250.0):
        self._ck(self.lib.btsdsp_rx_stream_host(self.h, _p(raw), nchunks, _p(tsc), nbursts, detect_thr, gate_thr, snr_thr,
                                                _p(flag), _p(amp), _p(toa), _p(soft), soft_pitch))

    def tx_stream_host(self, bits148, n, out):
        self._ck(self.lib.btsdsp_tx_stream_host(self.h, _p(bits148), n, _p(out)))

    def demod_normal_host(self, bursts, lens, tsc, detect_thr=3.0, gate_thr=-1.0, snr_thr=250.0, debug=True,
                          soft_pitch=160):
        """bursts (n, pitch) complex64 -> dict(flag, amp, toa, soft[, chan, off, w, b])"""
        bursts = _c64(bursts)
        n, pitch = bursts.shape
        lens = None if lens is None else np.ascontiguousarray(lens, np.int32)
        tsc = np.ascontiguousarray(tsc, np.uint8)
        r = dict(flag=np.zeros(n, np.int32), amp=np.zeros(n, np.complex64), toa=np.zeros(n, np.float32),
                 soft=np.zeros((n, soft_pitch), np.float32))
        if debug:
            r.update(chan=np.zeros((n, 6), np.complex64), off=np.zeros(n, np.float32),
                     w=np.zeros((n, 7), np.complex64), b=np.zeros((n, 5), np.complex64))
        self._ck(self.lib.btsdsp_demod_normal_host(
            self.h, _p(bursts), pitch, _p(lens), _p(tsc), n, detect_thr, gate_thr, snr_thr, _p(r["flag"]), _p(r["amp"]),
            _p(r["toa"]), _p(r["soft"]), soft_pitch, _p(r.get("chan")), _p(r.get("off")), _p(r.get("w")), _p(r.get("b"))))
        return r

    # ---- caller policy: Transceiver::pullRadioVector + driveReceiveFIFO over batches ----
    TRX_STATE_DTYPE = np.dtype([("thr", np.float64), ("prev_false_fn", np.int32), ("tsc", np.int32),
                                ("chan_type", np.int32, 8), ("est_fn", np.int32, 8), ("have", np.int32, 8),
                                ("snr", np.float32, 8), ("chan_off", np.float32, 8),
                                ("w", np.complex64, (8, 7)), ("b", np.complex64, (8, 5))], align=True)

    def trx_create(self, tsc, chan_type, start_fn=0):
        """tsc: (narfcn,) midamble codes; chan_type: (narfcn, 8) channel combinations.  Returns an opaque handle."""
        tsc = np.ascontiguousarray(tsc, np.uint8)
        ct = np.ascontiguousarray(chan_type, np.uint8).reshape(tsc.size, 8)
        h = ctypes.c_void_p()
        self._ck(self.lib.btsdsp_trx_create(self.h, tsc.size, _p(tsc), _p(ct), start_fn, ctypes.byref(h)))
        assert self.lib.btsdsp_trx_state_bytes() == self.TRX_STATE_DTYPE.itemsize
        return (h, tsc.size)

    def trx_destroy(self, trx):
        self._ck(self.lib.btsdsp_trx_destroy(self.h, trx[0]))

    def trx_set_slot(self, trx, arfcn, tn, chan_type):
        self._ck(self.lib.btsdsp_trx_set_slot(self.h, trx[0], arfcn, tn, chan_type))

    def trx_set_variant_52m(self, trx, enable=True, max_expected_delay=0):
        """the second transceiver variant's receive policy (Transceiver52M/Transceiver.cpp:268-404) for later pulls"""
        self._ck(self.lib.btsdsp_trx_set_variant_52m(self.h, trx[0], int(enable), int(max_expected_delay)))

    def trx_state(self, trx):
        st = np.zeros(trx[1], self.TRX_STATE_DTYPE)
        for a in range(trx[1]):
            self._ck(self.lib.btsdsp_trx_get_state(self.h, trx[0], a, _p(st[a:a + 1]), st.itemsize))
        return st

    def analyze_52m(self, burst, tsc, thr=3.0, max_toa=3, request=True):
        """Transceiver52M's analyzeTrafficBurst on one burst -> (detected, amp, toa, chan, off)"""
        burst = _c64(burst)
        det = ctypes.c_int(0)
        amp = np.zeros(1, np.complex64); toa = np.zeros(1, np.float32)
        chan = np.zeros(6 * self.sps, np.complex64); off = np.zeros(1, np.float32)
        self._ck(self.lib.btsdsp_analyze_traffic_burst_52m(self.h, _p(burst), burst.size, tsc, thr, max_toa, int(request),
                                                           ctypes.byref(det), _p(amp), _p(toa), _p(chan), _p(off)))
        return bool(det.value), amp[0], toa[0], chan, off[0]

    def analyze_52m_host(self, bursts, lens, tsc, thr=3.0, max_toa=3, request=True):
        """batched, through device buffers made here (torch-free): returns dict(flag, amp, toa, chan, off)"""
        import torch
        bursts = _c64(bursts)
        n, pitch = bursts.shape
        dev = torch.device("cuda:%d" % self.device)
        d_b = torch.from_numpy(bursts.view(np.float32).copy()).to(dev)
        d_l = torch.from_numpy(np.ascontiguousarray(lens, np.int32)).to(dev)
        d_t = torch.from_numpy(np.ascontiguousarray(tsc, np.uint8)).to(dev)
        flag = torch.zeros(n, dtype=torch.int32, device=dev); amp = torch.zeros(n * 2, device=dev)
        toa = torch.zeros(n, device=dev); chan = torch.zeros(n * 12 * self.sps, device=dev); off = torch.zeros(n, device=dev)
        self._ck(self.lib.btsdsp_analyze_52m_dev(self.h, _p(d_b), pitch, _p(d_l), 0, _p(d_t), n, thr, max_toa, int(request),
                                                 _p(flag), _p(amp), _p(toa), _p(chan), _p(off), None))
        torch.cuda.synchronize()
        return dict(flag=flag.cpu().numpy(), amp=amp.cpu().numpy().view(np.complex64), toa=toa.cpu().numpy(),
                    chan=chan.cpu().numpy().view(np.complex64).reshape(n, 6 * self.sps), off=off.cpu().numpy())

    def energy_detect_52m(self, v, win, thr):
        v = _c64(v)
        avg = ctypes.c_float(0); above = ctypes.c_int(0)
        self._ck(self.lib.btsdsp_energy_detect_52m(self.h, _p(v), v.size, win, thr, ctypes.byref(avg), ctypes.byref(above)))
        return bool(above.value), np.float32(avg.value)

    def xcch_decode_host(self, soft_u8):
        """soft_u8: (nframes*4, >=148) uint8 -> (u[nframes,228] uint8, ok[nframes] int32)"""
        soft_u8 = np.ascontiguousarray(soft_u8, np.uint8)
        n = soft_u8.shape[0] // 4
        u = np.zeros((n, 228), np.uint8)
        ok = np.zeros(n, np.int32)
        self._ck(self.lib.btsdsp_xcch_decode_host(self.h, _p(soft_u8), soft_u8.shape[1], n, _p(u), _p(ok)))
        return u, ok

    def rach_decode_host(self, soft_u8):
        """soft_u8: (n, >=148) uint8 -> (u[n,18], tail[n], bsic[n], ra[n])"""
        soft_u8 = np.ascontiguousarray(soft_u8, np.uint8)
        n = soft_u8.shape[0]
        u = np.zeros((n, 18), np.uint8)
        f = np.zeros(n, np.int32)
        self._ck(self.lib.btsdsp_rach_decode_host(self.h, _p(soft_u8), soft_u8.shape[1], n, _p(u), _p(f)))
        return u, f & 0xff, (f >> 8) & 0xff, (f >> 16) & 0xff

    # ---- CUDA graphs: record layer-2 calls on `stream`, replay them with one launch ----
    def graph_begin(self, stream):
        g = ctypes.c_void_p()
        self._ck(self.lib.btsdsp_graph_begin(self.h, _stream(stream), ctypes.byref(g)))
        return g

    def graph_end(self, graph, stream):
        self._ck(self.lib.btsdsp_graph_end(self.h, _stream(stream), graph))

    def graph_launch(self, graph, stream):
        self._ck(self.lib.btsdsp_graph_launch(self.h, graph, _stream(stream)))

    def graph_destroy(self, graph):
        self._ck(self.lib.btsdsp_graph_destroy(self.h, graph))

    def tch_decode_host(self, soft_u8):
        """soft_u8 (4*nblocks + 4, pitch >= 148) -> dict(d[nblocks,260], good, stolen, fu[nblocks,228], fok)"""
        soft_u8 = np.ascontiguousarray(soft_u8, np.uint8)
        n = soft_u8.shape[0] // 4 - 1
        r = dict(d=np.zeros((n, 260), np.uint8), good=np.zeros(n, np.int32), stolen=np.zeros(n, np.int32),
                 fu=np.zeros((n, 228), np.uint8), fok=np.zeros(n, np.int32))
        self._ck(self.lib.btsdsp_tch_decode_host(self.h, _p(soft_u8), soft_u8.shape[1], n, _p(r["d"]), _p(r["good"]), _p(r["stolen"]),
                                                 _p(r["fu"]), _p(r["fok"])))
        return r

    # ---- L1 encoders on the transmit side: L2 / speech frames -> 148-bit normal bursts ----
    def xcch_encode_host(self, frames, lsb8msb=True, tsc=-1, burst_pitch=148):
        """frames (n, 184) bits -> bursts (4n, burst_pitch) bits (XCCHL1Encoder::sendFrame)"""
        frames = np.ascontiguousarray(frames, np.uint8)
        n = frames.shape[0]
        out = np.zeros((4 * n, burst_pitch), np.uint8)
        self._ck(self.lib.btsdsp_xcch_encode_host(self.h, _p(frames), n, int(bool(lsb8msb)), tsc, _p(out), burst_pitch))
        return out

    def xcch_encode_dev(self, frames, nframes, lsb8msb, tsc, bursts, burst_pitch, stream=None):
        self._ck(self.lib.btsdsp_xcch_encode_dev(self.h, _p(frames), nframes, int(bool(lsb8msb)), tsc, _p(bursts), burst_pitch, _stream(stream)))

    def tch_encode_host(self, d260, f184, steal, lsb8msb=True, tsc=-1, carry=None, burst_pitch=148):
        """one traffic channel's blocks -> (4*nblocks + 4, burst_pitch) burst bits; the last four rows are the next call's carry"""
        d260 = np.ascontiguousarray(d260, np.uint8); f184 = np.ascontiguousarray(f184, np.uint8)
        steal = np.ascontiguousarray(steal, np.uint8)
        n = steal.shape[0]
        carry = None if carry is None else np.ascontiguousarray(carry, np.uint8)
        out = np.zeros((4 * n + 4, burst_pitch), np.uint8)
        self._ck(self.lib.btsdsp_tch_encode_host(self.h, _p(d260), _p(f184), _p(steal), n, int(bool(lsb8msb)), tsc,
                                                 None if carry is None else _p(carry), _p(out), burst_pitch))
        return out

    def tch_encode_dev(self, d260, f184, steal, nblocks, lsb8msb, tsc, carry, bursts, burst_pitch, stream=None):
        self._ck(self.lib.btsdsp_tch_encode_dev(self.h, _p(d260), _p(f184), _p(steal), nblocks, int(bool(lsb8msb)), tsc,
                                                None if carry is None else _p(carry), _p(bursts), burst_pitch, _stream(stream)))

    def tch_decode_dev(self, soft_u8, burst_pitch, nblocks, d, good, stolen, fu, fok, stream=None):
        self._ck(self.lib.btsdsp_tch_decode_dev(self.h, _p(soft_u8), burst_pitch, nblocks, _p(d), _p(good), _p(stolen), _p(fu), _p(fok),
                                                _stream(stream)))

    def xcch_decode_dev(self, soft_u8, burst_pitch, nframes, u, ok, stream=None):
        self._ck(self.lib.btsdsp_xcch_decode_dev(self.h, _p(soft_u8), burst_pitch, nframes, _p(u), _p(ok), _stream(stream)))

    def tx_datagrams_host(self, dgram, fn0, nframes, filler=None):
        """dgram: (n, >=154) uint8 TX datagrams.  Returns (iq[nchunks*864, 2] int16, placed)."""
        dgram = np.ascontiguousarray(dgram, np.uint8)
        filler = None if filler is None else np.ascontiguousarray(filler, np.uint8)
        nchunks = nframes * 1250 // 585
        out = np.zeros((nchunks * 864, 2), np.int16)
        placed = ctypes.c_longlong(0)
        self._ck(self.lib.btsdsp_tx_datagrams_host(self.h, _p(dgram), dgram.shape[0], dgram.shape[1], fn0, nframes,
                                                   _p(filler), _p(out), ctypes.byref(placed)))
        return out, placed.value

    def tx_datagrams_52m_host(self, dgram, fn0, nframes, filler=None):
        """the second variant's transmit side: (iq[nframes*1250, 2] int16 at the symbol rate, placed)"""
        dgram = np.ascontiguousarray(dgram, np.uint8)
        filler = None if filler is None else np.ascontiguousarray(filler, np.uint8)
        out = np.zeros((nframes * 1250, 2), np.int16)
        placed = ctypes.c_longlong(0)
        self._ck(self.lib.btsdsp_tx_datagrams_52m_host(self.h, _p(dgram), dgram.shape[0], dgram.shape[1], fn0, nframes,
                                                       _p(filler), _p(out), ctypes.byref(placed)))
        return out, placed.value

    def trx_pull_dev(self, trx, bursts, pitch, nframes, fn0, valid, dgram, dgram_pitch=160, stream=None):
        """device tensors/pointers; asynchronous on `stream`"""
        self._ck(self.lib.btsdsp_trx_pull_dev(self.h, trx[0], _p(bursts), pitch, nframes, fn0, _p(valid), _p(dgram),
                                              dgram_pitch, _stream(stream)))

    def trx_pull_streams_dev(self, trx, streams, stream_pitch, nframes, fn0, valid, dgram, dgram_pitch=160, stream=None):
        """per-ARFCN continuous slot streams (device), stream_pitch samples apart; asynchronous on `stream`"""
        self._ck(self.lib.btsdsp_trx_pull_streams_dev(self.h, trx[0], _p(streams), stream_pitch, nframes, fn0, _p(valid),
                                                      _p(dgram), dgram_pitch, _stream(stream)))

    def trx_radio_host(self, trx, iq, fn0, swap_iq=False):
        """iq: (narfcn, nchunks*864, 2) int16.  Returns (valid[n], dgram[n,158]) laid out [frame][arfcn][tn]."""
        iq = np.ascontiguousarray(iq, np.int16)
        A, ns, _ = iq.shape
        nchunks = ns // 864
        n = nchunks // 250 * 117 * A * 8
        valid = np.zeros(n, np.int32)
        dg = np.zeros((n, 158), np.uint8)
        self._ck(self.lib.btsdsp_trx_radio_host(self.h, trx[0], _p(iq), ns, nchunks, int(swap_iq), fn0, _p(valid), _p(dg), 158))
        return valid, dg

    def trx_pull_host(self, trx, bursts, fn0):
        """bursts: (nframes*narfcn*8, pitch) complex64 laid out [frame][arfcn][tn].  Returns (valid[n], dgram[n,158])."""
        bursts = _c64(bursts)
        n, pitch = bursts.shape
        assert n % (8 * trx[1]) == 0
        valid = np.zeros(n, np.int32)
        dg = np.zeros((n, 158), np.uint8)
        self._ck(self.lib.btsdsp_trx_pull_host(self.h, trx[0], _p(bursts), pitch, n // (8 * trx[1]), fn0, _p(valid), _p(dg), 158))
        return valid, dg

    def rach_host(self, bursts, lens, detect_thr=5.0, demod=True, soft_pitch=160):
        bursts = _c64(bursts)
        n, pitch = bursts.shape
        lens = None if lens is None else np.ascontiguousarray(lens, np.int32)
        r = dict(flag=np.zeros(n, np.int32), amp=np.zeros(n, np.complex64), toa=np.zeros(n, np.float32))
        if demod:
            r["soft"] = np.zeros((n, soft_pitch), np.float32)
        self._ck(self.lib.btsdsp_rach_host(self.h, _p(bursts), pitch, _p(lens), n, detect_thr, _p(r["flag"]),
                                           _p(r["amp"]), _p(r["toa"]), _p(r.get("soft")), soft_pitch))
        return r
