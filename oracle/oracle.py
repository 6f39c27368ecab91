"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

numpy/ctypes face over the two CPU checkers (same entry points, prefix ref_ / port_):

  Oracle("ref")   oracle/_ref/libref_oracle.so  -- the unmodified reference compiled in place
  Oracle("port")  oracle/libport_oracle.so      -- the plain-C restatement (oracle/sigproc_port.c)
  Oracle("best")  ref when present, else port

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (openbts_ttsou_b200) never does.
Complex vectors are numpy complex64; bits uint8/int8 (one per bit); soft bits float32.
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libref_oracle.so")
PORT_SO = os.path.join(HERE, "libport_oracle.so")

c_f = ctypes.c_float
c_i = ctypes.c_int
c_l = ctypes.c_long
c_p = ctypes.c_void_p

FULL_SPAN, OVERLAP_ONLY, START_ONLY, WITH_TAIL, NO_DELAY = 0, 1, 2, 3, 4
T_COS, T_SIN, T_ROT, T_REVROT, T_PULSE, T_MID_SEQ, T_MID_META, T_RACH_SEQ, T_RACH_META, T_LPF_RX, T_LPF_TX, T_OOB = range(12)


def build(ref=True):
    """(Re)build the checkers with oracle/Makefile.  ref is skipped by make when /root/reference is absent."""
    subprocess.run(["make", "-C", HERE, "port"] + (["ref", "ref52"] if ref else []), check=True,
                   stdout=subprocess.DEVNULL)


def have_ref():
    return os.path.exists(REF_SO)


def _ptr(a):
    return a.ctypes.data_as(c_p)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.complex64)


class Oracle52:
    """The reference's second transceiver variant (Transceiver52M/sigProcLib.cpp) compiled in place: only the functions
    that differ from the main variant.  Exists only where oracle/_ref/libref52_oracle.so was built."""
    SO = os.path.join(HERE, "_ref", "libref52_oracle.so")

    def __init__(self, sps=1):
        if not os.path.exists(self.SO):
            raise FileNotFoundError(self.SO)
        self.lib = ctypes.CDLL(self.SO)
        self.sps = sps
        self.lib.ref52_setup(c_i(sps))

    def analyze(self, burst, tsc, thr=3.0, max_toa=3, request=True):
        burst = _c64(burst)
        amp = np.zeros(1, np.complex64); toa = np.zeros(1, np.float32)
        chan = np.zeros(6 * self.sps, np.complex64); off = np.zeros(1, np.float32)
        self.lib.ref52_setup(c_i(self.sps))
        ok = self.lib.ref52_analyze(_ptr(burst), c_i(burst.size), c_i(tsc), c_f(thr), c_i(self.sps), ctypes.c_uint(max_toa),
                                    c_i(int(request)), _ptr(amp), _ptr(toa), _ptr(chan), _ptr(off))
        return bool(ok), amp[0], toa[0], chan, off[0]

    # ---- that variant's caller policy and transmit side (Transceiver52M/Transceiver.cpp) ----
    def trx_new(self, tsc, chan_type, start_fn=0):
        self.lib.ref52_setup(c_i(1))
        assert self.lib.ref52_trx_state_bytes() == Oracle.TRX_STATE_DTYPE.itemsize
        st = np.zeros(1, Oracle.TRX_STATE_DTYPE)
        ct = np.ascontiguousarray(chan_type, np.int32)
        self.lib.ref52_trx_init(_ptr(st), c_i(tsc), _ptr(ct), c_i(start_fn))
        return st

    def trx_pull(self, st, bursts, fn0, max_expected_delay):
        """bursts: (nframes*8, pitch) complex64 of one ARFCN in FIFO order.  Returns (valid[n], dgram[n,160])."""
        bursts = _c64(bursts)
        n, pitch = bursts.shape
        valid = np.zeros(n, np.int32)
        dg = np.zeros((n, 160), np.uint8)
        self.lib.ref52_setup(c_i(1))
        self.lib.ref52_trx_pull(_ptr(st), _ptr(bursts), c_i(pitch), c_i(n // 8), c_i(fn0), c_i(max_expected_delay), _ptr(valid),
                                _ptr(dg), c_i(160))
        return valid, dg

    def tx_datagrams(self, dgram, fn0, nframes, filler=None):
        """TX datagrams (n, >=154) uint8 -> (iq[nframes*1250, 2] int16 at the symbol rate, placed)"""
        dgram = np.ascontiguousarray(dgram, np.uint8)
        filler = None if filler is None else np.ascontiguousarray(filler, np.uint8)
        out = np.zeros((nframes * 1250, 2), np.int16)
        self.lib.ref52_setup(c_i(1))
        self.lib.ref52_tx_datagrams.restype = ctypes.c_long
        placed = self.lib.ref52_tx_datagrams(_ptr(dgram), c_l(dgram.shape[0]), c_i(dgram.shape[1]), c_i(fn0), c_i(nframes),
                                             None if filler is None else _ptr(filler), _ptr(out))
        return out, placed

    def energy_detect(self, v, win, thr):
        v = _c64(v)
        avg = np.zeros(1, np.float32)
        ok = self.lib.ref52_energy_detect(_ptr(v), c_i(v.size), ctypes.c_uint(win), c_f(thr), _ptr(avg))
        return bool(ok), avg[0]


class Oracle:
    def __init__(self, kind="best", sps=1):
        if kind == "best":
            kind = "ref" if have_ref() else "port"
        self.kind = kind
        path = REF_SO if kind == "ref" else PORT_SO
        if not os.path.exists(path):
            if kind == "port":
                build(ref=False)
            else:
                raise FileNotFoundError(path)
        self.lib = ctypes.CDLL(path)
        self.pfx = "ref_" if kind == "ref" else "port_"
        for name in ("sinc", "sin_lookup", "cos_lookup"):
            f = self._f(name)
            f.restype = c_f
            f.argtypes = [c_f]
        self._f("modulate_stream").restype = c_l
        self.sps = None
        self.setup(sps)

    def _f(self, name):
        return getattr(self.lib, self.pfx + name)

    def setup(self, sps):
        if self.sps != sps:
            self._f("setup")(c_i(sps))
            self.sps = sps

    # ---- tables
    def table(self, tid, idx=0):
        buf = np.zeros(4096, np.float32)
        n = self._f("get_table")(c_i(tid), c_i(idx), _ptr(buf), c_i(buf.size))
        assert n >= 0
        out = buf[:n].copy()
        if tid in (T_ROT, T_REVROT, T_PULSE, T_MID_SEQ, T_RACH_SEQ):
            return out.view(np.complex64)
        return out

    def sinc(self, x):
        return float(self._f("sinc")(c_f(x)))

    def sin_lookup(self, x):
        return float(self._f("sin_lookup")(c_f(x)))

    def cos_lookup(self, x):
        return float(self._f("cos_lookup")(c_f(x)))

    # ---- single-vector functions (sigProcLib.h surface)
    def modulate(self, bits, guard, sps=None):
        sps = sps or self.sps
        bits = np.ascontiguousarray(bits, dtype=np.int8)
        out = np.zeros(sps * (bits.size + guard), np.complex64)
        n = self._f("modulate")(_ptr(bits), c_i(bits.size), c_i(guard), c_i(sps), _ptr(out), c_i(out.size))
        assert n == out.size
        return out

    def delay_vector(self, v, delay):
        v = _c64(v).copy()
        self._f("delay_vector")(_ptr(v), c_i(v.size), c_f(delay))
        return v

    def scale_vector(self, v, scale, real_only=False):
        v = _c64(v).copy()
        s = np.array([complex(scale)], np.complex64)
        self._f("scale_vector")(_ptr(v), c_i(v.size), c_i(int(real_only)), _ptr(s))
        return v

    def _conv(self, name, a, b, span, a_real, b_real):
        a, b = _c64(a), _c64(b)
        out = np.zeros(a.size + b.size + 2, np.complex64)
        n = self._f(name)(_ptr(a), c_i(a.size), c_i(int(a_real)), _ptr(b), c_i(b.size), c_i(int(b_real)),
                          _ptr(out), c_i(out.size), c_i(span))
        assert n >= 0
        return out[:n].copy()

    def convolve(self, a, b, span, a_real=False, b_real=False):
        return self._conv("convolve", a, b, span, a_real, b_real)

    def correlate(self, a, b, span, a_real=False, b_real=False):
        return self._conv("correlate", a, b, span, a_real, b_real)

    def peak_detect(self, v):
        v = _c64(v)
        pk = np.zeros(1, np.complex64)
        idx, avg = c_f(), c_f()
        self._f("peak_detect")(_ptr(v), c_i(v.size), _ptr(pk), ctypes.byref(idx), ctypes.byref(avg))
        return pk[0], idx.value, avg.value

    def interpolate_point(self, v, ix):
        v = _c64(v)
        pk = np.zeros(1, np.complex64)
        self._f("interpolate_point")(_ptr(v), c_i(v.size), c_f(ix), _ptr(pk))
        return pk[0]

    def energy_detect(self, v, win, thr):
        v = _c64(v)
        avg = c_f()
        ok = self._f("energy_detect")(_ptr(v), c_i(v.size), ctypes.c_uint(win), c_f(thr), ctypes.byref(avg))
        return bool(ok), avg.value

    def analyze(self, burst, tsc, thr, sps=None, request=True):
        sps = sps or self.sps
        burst = _c64(burst)
        amp = np.zeros(1, np.complex64)
        chan = np.zeros(6 * sps, np.complex64)
        toa, off = c_f(), c_f()
        ok = self._f("analyze")(_ptr(burst), c_i(burst.size), c_i(tsc), c_f(thr), c_i(sps), _ptr(amp),
                                ctypes.byref(toa), c_i(int(request)), _ptr(chan), ctypes.byref(off))
        return bool(ok), amp[0], toa.value, chan, off.value

    def detect_rach(self, burst, thr, sps=None):
        sps = sps or self.sps
        burst = _c64(burst)
        amp = np.zeros(1, np.complex64)
        toa = c_f()
        ok = self._f("detect_rach")(_ptr(burst), c_i(burst.size), c_f(thr), c_i(sps), _ptr(amp), ctypes.byref(toa))
        return bool(ok), amp[0], toa.value

    def design_dfe(self, chan, snr, nf=7):
        chan = _c64(chan)
        w = np.zeros(nf, np.complex64)
        b = np.zeros(chan.size - 1, np.complex64)
        self._f("design_dfe")(_ptr(chan), c_i(chan.size), c_f(snr), c_i(nf), _ptr(w), _ptr(b))
        return w, b

    def equalize(self, burst, toa, w, b, sps=1):
        """returns (soft, burst_after) -- the reference delays the burst in place"""
        burst = _c64(burst).copy()
        w, b = _c64(w), _c64(b)
        soft = np.zeros(burst.size, np.float32)
        self._f("equalize")(_ptr(burst), c_i(burst.size), c_f(toa), c_i(sps), _ptr(w), c_i(w.size), _ptr(b),
                            c_i(b.size), _ptr(soft))
        return soft, burst

    def demodulate(self, burst, amp, toa, sps=None):
        sps = sps or self.sps
        burst = _c64(burst)
        a = np.array([amp], np.complex64)
        soft = np.zeros(burst.size, np.float32)
        n = self._f("demodulate")(_ptr(burst), c_i(burst.size), c_i(sps), _ptr(a), c_f(toa), _ptr(soft))
        return soft[:n].copy()

    def resample(self, x, P, Q, lpf):
        x = _c64(x)
        out = np.zeros(int(np.ceil(x.size * P / Q)) + 4, np.complex64)
        n = self._f("resample")(_ptr(x), c_i(x.size), c_i(P), c_i(Q), c_i(lpf), _ptr(out), c_i(out.size))
        return out[:n].copy()

    # ---- batched caller glue; `threads` python threads each run a contiguous slice (ctypes drops the GIL)
    @staticmethod
    def _shard(n, threads):
        threads = max(1, min(threads, n))
        edges = np.linspace(0, n, threads + 1).astype(np.int64)
        return [(int(edges[i]), int(edges[i + 1])) for i in range(threads) if edges[i + 1] > edges[i]]

    @staticmethod
    def _run(jobs, fn):
        if len(jobs) == 1:
            fn(*jobs[0])
            return
        with ThreadPoolExecutor(len(jobs)) as ex:
            list(ex.map(lambda j: fn(*j), jobs))

    def rx_normal_batch(self, bursts, lens, tsc, detect_thr=3.0, energy_thr=250.0, threads=1, debug=True):
        """bursts: (n, pitch) complex64.  Returns dict(flag, amp, toa, chan, off, w, b, soft[n,160])."""
        bursts = _c64(bursts)
        n, pitch = bursts.shape
        lens = np.ascontiguousarray(lens, np.int32)
        tsc = np.ascontiguousarray(tsc, np.uint8)
        r = dict(flag=np.zeros(n, np.int32), amp=np.zeros(n, np.complex64), toa=np.zeros(n, np.float32),
                 soft=np.zeros((n, 160), np.float32))
        if debug:
            r.update(chan=np.zeros((n, 6), np.complex64), off=np.zeros(n, np.float32),
                     w=np.zeros((n, 7), np.complex64), b=np.zeros((n, 5), np.complex64))
        f = self._f("rx_normal_batch")

        def job(lo, hi):
            f(_ptr(bursts[lo:hi]), c_i(pitch), _ptr(lens[lo:hi]), _ptr(tsc[lo:hi]), c_l(hi - lo),
              c_f(detect_thr), c_f(energy_thr), _ptr(r["flag"][lo:hi]), _ptr(r["amp"][lo:hi]), _ptr(r["toa"][lo:hi]),
              _ptr(r["chan"][lo:hi]) if debug else None, _ptr(r["off"][lo:hi]) if debug else None,
              _ptr(r["w"][lo:hi]) if debug else None, _ptr(r["b"][lo:hi]) if debug else None,
              _ptr(r["soft"][lo:hi]), c_i(160))
        self._run(self._shard(n, threads), job)
        return r

    # ---- Transceiver::pullRadioVector policy + RX datagram ----
    TRX_STATE_DTYPE = np.dtype([("thr", np.float64), ("prev_false_fn", np.int32), ("tsc", np.int32),
                                ("chan_type", np.int32, 8), ("est_fn", np.int32, 8), ("have", np.int32, 8),
                                ("snr", np.float32, 8), ("chan_off", np.float32, 8),
                                ("w", np.complex64, (8, 7)), ("b", np.complex64, (8, 5))], align=True)

    def expected_corr_type(self, chan_type, fn):
        f = self._f("expected_corr_type")
        f.restype = ctypes.c_int
        return f(c_i(chan_type), c_i(fn))

    def trx_new(self, tsc, chan_type, start_fn=0):
        """one Transceiver object's receive-side state (one ARFCN)"""
        f = self._f("trx_state_bytes")
        f.restype = ctypes.c_int
        assert f() == self.TRX_STATE_DTYPE.itemsize, (f(), self.TRX_STATE_DTYPE.itemsize)
        st = np.zeros(1, self.TRX_STATE_DTYPE)
        ct = np.ascontiguousarray(chan_type, np.int32)
        self._f("trx_init")(_ptr(st), c_i(tsc), _ptr(ct), c_i(start_fn))
        return st

    def trx_pull(self, st, bursts, fn0):
        """bursts: (nframes*8, pitch) complex64 in FIFO order (frame-major, TN 0..7).  Updates st in place.
        Returns (valid[n] int32, dgram[n,160] uint8: 158 datagram bytes + 2 pad)."""
        bursts = _c64(bursts)
        n, pitch = bursts.shape
        assert n % 8 == 0
        valid = np.zeros(n, np.int32)
        dg = np.zeros((n, 160), np.uint8)
        self._f("trx_pull")(_ptr(st), _ptr(bursts), c_i(pitch), c_i(n // 8), c_i(fn0), _ptr(valid), _ptr(dg), c_i(160))
        return valid, dg

    def tx_datagrams(self, dgram, fn0, nframes, filler=None):
        """TX datagrams (n, >=154) uint8 -> (iq[nchunks*864, 2] int16, placed); reference glue, see ref_shim.cpp"""
        assert self.kind == "ref", "only the compiled reference implements the TX datagram glue"
        dgram = np.ascontiguousarray(dgram, np.uint8)
        filler = None if filler is None else np.ascontiguousarray(filler, np.uint8)
        nchunks = nframes * 1250 // 585
        out = np.zeros((nchunks * 864, 2), np.int16)
        f = self._f("tx_datagrams")
        f.restype = ctypes.c_long
        placed = f(_ptr(dgram), c_l(dgram.shape[0]), c_i(dgram.shape[1]), c_i(fn0), c_i(nframes),
                   None if filler is None else _ptr(filler), _ptr(out))
        return out, placed

    # ---- L1 FEC after the path: XCCH deinterleave + Viterbi + parity ----
    def xcch_decode(self, soft_u8):
        """soft_u8: (nframes*4, >=148) uint8 soft bytes of consecutive bursts.  Returns (u[nframes,228] uint8, ok[nframes])."""
        soft_u8 = np.ascontiguousarray(soft_u8, np.uint8)
        n = soft_u8.shape[0] // 4
        u = np.zeros((n, 228), np.uint8)
        ok = np.zeros(n, np.int32)
        self._f("xcch_decode")(_ptr(soft_u8), c_i(soft_u8.shape[1]), c_l(n), _ptr(u), _ptr(ok))
        return u, ok

    def xcch_encode(self, d):
        """d: (nframes, 184) bits -> e-bits (nframes*4, 114) of the four bursts (reference encoder; ref only)"""
        assert self.kind == "ref"
        d = np.ascontiguousarray(d, np.uint8)
        e = np.zeros((d.shape[0] * 4, 114), np.uint8)
        self._f("xcch_encode")(_ptr(d), c_l(d.shape[0]), _ptr(e))
        return e

    def xcch_encode_pflip(self, d, pflip):
        """xcch_encode with the 40-bit parity word of frame f XORed by pflip[f] (ref only)"""
        assert self.kind == "ref"
        d = np.ascontiguousarray(d, np.uint8)
        pf = np.ascontiguousarray(pflip, np.uint64)
        e = np.zeros((d.shape[0] * 4, 114), np.uint8)
        self._f("xcch_encode_pflip")(_ptr(d), _ptr(pf), c_l(d.shape[0]), _ptr(e))
        return e

    def tch_decode(self, soft_u8):
        """soft_u8: (4*nblocks + 4, >=148) soft bytes of one traffic channel's consecutive bursts.
        Returns dict(d[nblocks,260], good, stolen, fu[nblocks,228], fok)."""
        soft_u8 = np.ascontiguousarray(soft_u8, np.uint8)
        n = soft_u8.shape[0] // 4 - 1
        r = dict(d=np.zeros((n, 260), np.uint8), good=np.zeros(n, np.int32), stolen=np.zeros(n, np.int32),
                 fu=np.zeros((n, 228), np.uint8), fok=np.zeros(n, np.int32))
        self._f("tch_decode")(_ptr(soft_u8), c_i(soft_u8.shape[1]), c_l(n), _ptr(r["d"]), _ptr(r["good"]), _ptr(r["stolen"]),
                              _ptr(r["fu"]), _ptr(r["fok"]))
        return r

    def tch_encode(self, d260, f184, steal):
        """nblocks speech frames d260 / FACCH payloads f184 (used where steal) -> burst bits (4*nblocks + 4, 148) (ref only)"""
        assert self.kind == "ref"
        d260 = np.ascontiguousarray(d260, np.uint8); f184 = np.ascontiguousarray(f184, np.uint8)
        steal = np.ascontiguousarray(steal, np.uint8)
        n = d260.shape[0]
        out = np.zeros((4 * n + 4, 148), np.uint8)
        self._f("tch_encode")(_ptr(d260), _ptr(f184), _ptr(steal), c_l(n), _ptr(out))
        return out

    # ---- L1 encoders on the transmit side, the reference's own flow (ref only) ----
    def xcch_send_frames(self, frames, lsb8msb=True, tsc=-1):
        """frames (n, 184) bits -> (4n, 148) burst bits: XCCHL1Encoder::sendFrame .. transmit"""
        assert self.kind == "ref"
        frames = np.ascontiguousarray(frames, np.uint8)
        out = np.zeros((4 * frames.shape[0], 148), np.uint8)
        self._f("xcch_send_frames")(_ptr(frames), c_l(frames.shape[0]), c_i(int(bool(lsb8msb))), c_i(tsc), _ptr(out))
        return out

    def tch_dispatch(self, d260, f184, steal, lsb8msb=True, tsc=-1, state=None):
        """one TCHFACCHL1Encoder::dispatch per block -> ((4*nblocks, 148) burst bits, state); state = the encoder's members
        (interleaver rows still to be sent, mPreviousFACCH, mOffset) carried to the next call, None = a fresh encoder"""
        assert self.kind == "ref"
        d260 = np.ascontiguousarray(d260, np.uint8); f184 = np.ascontiguousarray(f184, np.uint8)
        steal = np.ascontiguousarray(steal, np.uint8)
        n = steal.shape[0]
        state = np.zeros(458, np.uint8) if state is None else np.ascontiguousarray(state, np.uint8).copy()
        out = np.zeros((4 * n, 148), np.uint8)
        self._f("tch_dispatch")(_ptr(d260), _ptr(f184), _ptr(steal), c_l(n), c_i(int(bool(lsb8msb))), c_i(tsc), _ptr(state), _ptr(out))
        return out, state

    def rach_decode(self, soft_u8):
        """soft_u8: (n, >=148) uint8 -> (u[n,18], tail[n], bsic[n], ra[n]); reference classes (ref only)"""
        assert self.kind == "ref"
        soft_u8 = np.ascontiguousarray(soft_u8, np.uint8)
        n = soft_u8.shape[0]
        u = np.zeros((n, 18), np.uint8)
        tail, bsic, ra = (np.zeros(n, np.int32) for _ in range(3))
        self._f("rach_decode")(_ptr(soft_u8), c_i(soft_u8.shape[1]), c_l(n), _ptr(u), _ptr(tail), _ptr(bsic), _ptr(ra))
        return u, tail, bsic, ra

    def rach_encode(self, ra, bsic):
        assert self.kind == "ref"
        ra = np.ascontiguousarray(ra, np.uint8)
        bsic = np.ascontiguousarray(bsic, np.uint8)
        e = np.zeros((ra.size, 36), np.uint8)
        self._f("rach_encode")(_ptr(ra), _ptr(bsic), c_l(ra.size), _ptr(e))
        return e

    def rx_rach_batch(self, bursts, lens, detect_thr=5.0, sps=1, threads=1):
        bursts = _c64(bursts)
        n, pitch = bursts.shape
        lens = np.ascontiguousarray(lens, np.int32)
        spitch = 160
        r = dict(flag=np.zeros(n, np.int32), amp=np.zeros(n, np.complex64), toa=np.zeros(n, np.float32),
                 soft=np.zeros((n, spitch), np.float32))
        f = self._f("rx_rach_batch")

        def job(lo, hi):
            f(_ptr(bursts[lo:hi]), c_i(pitch), _ptr(lens[lo:hi]), c_l(hi - lo), c_f(detect_thr), c_i(sps),
              _ptr(r["flag"][lo:hi]), _ptr(r["amp"][lo:hi]), _ptr(r["toa"][lo:hi]), _ptr(r["soft"][lo:hi]), c_i(spitch))
        self._run(self._shard(n, threads), job)
        return r

    def rx_resample_stream(self, raw, threads=1):
        """raw: complex64, length multiple of 864, stream starts at chunk 0 (zero history)."""
        raw = _c64(raw)
        nch = raw.size // 864
        out = np.zeros(nch * 585, np.complex64)
        f = self._f("rx_resample_stream")

        def job(lo, hi):
            f(_ptr(raw[lo * 864:]), c_l(lo), c_l(hi - lo), _ptr(out[lo * 585:]))
        self._run(self._shard(nch, threads), job)
        return out

    def rx_resample_stream_i16(self, iq, flip_iq=False, threads=1, out=None):
        """iq: int16 (n, 2) radio samples, n a multiple of 864, the stream starts at chunk 0 -> complex64 at 1 sps
        (unUSRPifyVector + pullBuffer's resample, radioInterface.cpp:91-116, 238-259).  out: optional preallocated result."""
        iq = np.ascontiguousarray(iq, np.int16).reshape(-1, 2)
        nch = iq.shape[0] // 864
        if out is None:
            out = np.zeros(nch * 585, np.complex64)
        assert out.dtype == np.complex64 and out.size >= nch * 585
        f = self._f("rx_resample_stream_i16")

        def job(lo, hi):
            f(_ptr(iq[lo * 864:]), c_i(1 if flip_iq else 0), c_l(lo), c_l(hi - lo), _ptr(out[lo * 585:]))
        self._run(self._shard(nch, threads), job)
        return out

    def soft_to_wire(self, soft, threads=1, out=None):
        """soft (n, pitch >= 148) float32 -> (n, 148) uint8 as the RX datagram carries them (Transceiver.cpp:667-669)"""
        soft = np.ascontiguousarray(soft, np.float32)
        n, pitch = soft.shape
        if out is None:
            out = np.zeros((n, 148), np.uint8)
        f = self._f("soft_to_wire")

        def job(lo, hi):
            f(_ptr(soft[lo:hi]), c_i(pitch), c_l(hi - lo), _ptr(out[lo:hi]))
        self._run(self._shard(n, threads), job)
        return out

    def tx_resample_stream(self, x, threads=1):
        """x: complex64 at 1 sps, length multiple of 585 -> int16 (n,2) at 400 kS/s, 864 per chunk."""
        x = _c64(x)
        nch = x.size // 585
        out = np.zeros((nch * 864, 2), np.int16)
        f = self._f("tx_resample_stream")

        def job(lo, hi):
            f(_ptr(x[lo * 585:]), c_l(lo), c_l(hi - lo), _ptr(out[lo * 864:]))
        self._run(self._shard(nch, threads), job)
        return out

    def modulate_stream(self, bits148, tn0=0, threads=1):
        """bits148: (n,148) -> complex64 stream of 157/156/156/156-sample bursts (n multiple of 4 per shard)."""
        bits = np.ascontiguousarray(bits148, np.int8)
        n = bits.shape[0]
        out = np.zeros(stream_offset(n, tn0) + 0, np.complex64)
        f = self._f("modulate_stream")
        f.restype = c_l

        def job(lo, hi):
            f(_ptr(bits[lo:hi]), c_l(hi - lo), c_i((tn0 + lo) % 8), _ptr(out[stream_offset(lo, tn0):]))
        jobs = [(lo - lo % 8, hi - hi % 8 if hi < n else hi) for lo, hi in self._shard(n, threads)]
        self._run([j for j in jobs if j[1] > j[0]], job)
        return out

    def rx_stream_demod(self, resampled, nbursts, tsc, detect_thr=3.0, energy_thr=250.0, threads=1, out=None):
        res = _c64(resampled)
        tsc = np.ascontiguousarray(tsc, np.uint8)
        assert stream_offset(nbursts) <= res.size
        r = out if out is not None else dict(flag=np.zeros(nbursts, np.int32), amp=np.zeros(nbursts, np.complex64),
                                             toa=np.zeros(nbursts, np.float32), soft=np.zeros((nbursts, 160), np.float32))
        f = self._f("rx_stream_demod")

        def job(lo, hi):
            f(_ptr(res), c_l(lo), c_l(hi - lo), _ptr(tsc[lo:hi]), c_f(detect_thr), c_f(energy_thr),
              _ptr(r["flag"][lo:hi]), _ptr(r["amp"][lo:hi]), _ptr(r["toa"][lo:hi]), _ptr(r["soft"][lo:hi]), c_i(160))
        self._run(self._shard(nbursts, threads), job)
        return r


def stream_offset(burst, tn0=0):
    """sample offset of burst `burst` in a 157/156/156/156 slot stream that starts at timeslot tn0 (tn0 % 4 == 0)."""
    assert tn0 % 4 == 0
    return (burst // 4) * 625 + (0, 157, 313, 469)[burst % 4]
