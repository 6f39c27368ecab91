"""numpy face over tests/hostemu/libhostemu.so (the kernels' bodies compiled for the CPU); test infrastructure."""
import ctypes

import numpy as np

c_p, c_i, c_ll, c_f = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float


def P(a):
    return None if a is None else a.ctypes.data_as(c_p)


class Emu:
    def __init__(self, lib, sps=1):
        self.lib = lib
        self.sps = None
        self.setup(sps)

    def setup(self, sps):
        if sps != self.sps:
            self.lib.emu_setup(sps)
            self.sps = sps

    def table(self, tid, idx=0):
        buf = np.zeros(4096, np.float32)
        n = self.lib.emu_get_table(tid, idx, P(buf), 4096)
        out = buf[:n].copy()
        return out.view(np.complex64) if tid in (2, 3, 4, 5, 7) else out

    def rx_normal_batch(self, bursts, lens, tsc, detect_thr=3.0, gate_thr=-1.0, snr_thr=250.0, first=0, pitch=None):
        bursts = np.ascontiguousarray(bursts, np.complex64)
        n = len(tsc)
        pitch = bursts.shape[1] if pitch is None else pitch
        lens = None if lens is None else np.ascontiguousarray(lens, np.int32)
        tsc = np.ascontiguousarray(tsc, np.uint8)
        r = dict(flag=np.zeros(n, np.int32), amp=np.zeros(n, np.complex64), toa=np.zeros(n, np.float32),
                 soft=np.zeros((n, 160), np.float32), chan=np.zeros((n, 6), np.complex64), off=np.zeros(n, np.float32),
                 w=np.zeros((n, 7), np.complex64), b=np.zeros((n, 5), np.complex64))
        self.lib.emu_demod_normal(P(bursts), c_ll(pitch), P(lens), c_ll(first), P(tsc), c_ll(n), c_f(detect_thr),
                                  c_f(gate_thr), c_f(snr_thr), P(r["flag"]), P(r["amp"]), P(r["toa"]), P(r["soft"]),
                                  c_i(160), P(r["chan"]), P(r["off"]), P(r["w"]), P(r["b"]))
        return r

    # ---- caller policy pipeline (trx_policy.cuh) ----
    def trx_new(self, tsc, chan_type, start_fn=0):
        from oracle.oracle import Oracle
        tsc = np.ascontiguousarray(tsc, np.uint8)
        ct = np.ascontiguousarray(chan_type, np.uint8).reshape(tsc.size, 8)
        st = np.zeros(tsc.size, Oracle.TRX_STATE_DTYPE)
        assert self.lib.emu_trx_state_bytes() == st.itemsize
        self.lib.emu_trx_init(P(st), c_i(tsc.size), P(tsc), P(ct), c_i(start_fn))
        return st

    def trx_pull(self, st, bursts, fn0):
        bursts = np.ascontiguousarray(bursts, np.complex64)
        n, pitch = bursts.shape
        valid = np.zeros(n, np.int32)
        dg = np.zeros((n, 158), np.uint8)
        self.lib.emu_trx_pull(P(st), c_i(st.size), P(bursts), c_ll(pitch), c_i(n // (8 * st.size)), c_i(fn0), P(valid),
                              P(dg), c_i(158))
        return valid, dg

    def trx_pull_52m(self, st, bursts, fn0, max_delay):
        bursts = np.ascontiguousarray(bursts, np.complex64)
        n, pitch = bursts.shape
        valid = np.zeros(n, np.int32)
        dg = np.zeros((n, 158), np.uint8)
        self.lib.emu_trx_pull_52m(P(st), c_i(st.size), P(bursts), c_ll(pitch), c_i(n // (8 * st.size)), c_i(fn0), c_i(max_delay),
                                  P(valid), P(dg), c_i(158))
        return valid, dg

    def tx_fused(self, bits, scale=None):
        """bits (nslots, 148) uint8 -> int16 (nchunks*864, 2): the fused TX kernel's arithmetic on the CPU"""
        bits = np.ascontiguousarray(bits, np.uint8)
        nslots = bits.shape[0]
        scale = None if scale is None else np.ascontiguousarray(scale, np.float32)
        out = np.zeros((nslots // 4 * 625 // 585 * 864, 2), np.int16)
        self.lib.emu_tx_fused(P(bits), P(scale), c_ll(nslots), P(out))
        return out

    def tch_decode(self, soft_u8):
        soft_u8 = np.ascontiguousarray(soft_u8, np.uint8)
        n = soft_u8.shape[0] // 4 - 1
        r = dict(d=np.zeros((n, 260), np.uint8), good=np.zeros(n, np.int32), stolen=np.zeros(n, np.int32),
                 fu=np.zeros((n, 228), np.uint8), fok=np.zeros(n, np.int32))
        self.lib.emu_tch_decode(P(soft_u8), c_i(soft_u8.shape[1]), c_ll(n), P(r["d"]), P(r["good"]), P(r["stolen"]), P(r["fu"]), P(r["fok"]))
        return r

    def xcch_encode(self, frames, lsb8msb=True, tsc=-1, burst_pitch=148):
        frames = np.ascontiguousarray(frames, np.uint8)
        out = np.zeros((4 * frames.shape[0], burst_pitch), np.uint8)
        self.lib.emu_xcch_encode(P(frames), c_ll(frames.shape[0]), c_i(int(bool(lsb8msb))), c_i(tsc), P(out), c_i(burst_pitch))
        return out

    def xcch_encode_lanes(self, frames, lsb8msb=True, tsc=-1, popc=False):
        frames = np.ascontiguousarray(frames, np.uint8)
        out = np.zeros((4 * frames.shape[0], 148), np.uint8)
        (self.lib.emu_xcch_encode_lanes_popc if popc else self.lib.emu_xcch_encode_lanes)(P(frames), c_ll(frames.shape[0]), c_i(int(bool(lsb8msb))), c_i(tsc), P(out))
        return out

    def tch_encode_lanes(self, d260, f184, steal, lsb8msb=True, tsc=-1):
        """lane form: rows 4 .. of the (4*nblocks + 4, 148) output (groups 1 .. nblocks); rows 0..3 stay zero"""
        d260 = np.ascontiguousarray(d260, np.uint8); f184 = np.ascontiguousarray(f184, np.uint8)
        steal = np.ascontiguousarray(steal, np.uint8)
        out = np.zeros((4 * steal.shape[0] + 4, 148), np.uint8)
        self.lib.emu_tch_encode_lanes(P(d260), P(f184), P(steal), c_ll(steal.shape[0]), c_i(int(bool(lsb8msb))), c_i(tsc), P(out))
        return out

    def tch_encode(self, d260, f184, steal, lsb8msb=True, tsc=-1, carry=None, burst_pitch=148):
        d260 = np.ascontiguousarray(d260, np.uint8); f184 = np.ascontiguousarray(f184, np.uint8)
        steal = np.ascontiguousarray(steal, np.uint8)
        carry = None if carry is None else np.ascontiguousarray(carry, np.uint8)
        out = np.zeros((4 * steal.shape[0] + 4, burst_pitch), np.uint8)
        self.lib.emu_tch_encode(P(d260), P(f184), P(steal), c_ll(steal.shape[0]), c_i(int(bool(lsb8msb))), c_i(tsc),
                                None if carry is None else P(carry), P(out), c_i(burst_pitch))
        return out

    def xcch_decode(self, soft_u8):
        soft_u8 = np.ascontiguousarray(soft_u8, np.uint8)
        n = soft_u8.shape[0] // 4
        u = np.zeros((n, 228), np.uint8)
        ok = np.zeros(n, np.int32)
        self.lib.emu_xcch_decode(P(soft_u8), c_i(soft_u8.shape[1]), c_ll(n), P(u), P(ok))
        return u, ok

    def rach_decode(self, soft_u8):
        soft_u8 = np.ascontiguousarray(soft_u8, np.uint8)
        n = soft_u8.shape[0]
        u = np.zeros((n, 18), np.uint8)
        tail, bsic, ra = (np.zeros(n, np.int32) for _ in range(3))
        self.lib.emu_rach_decode(P(soft_u8), c_i(soft_u8.shape[1]), c_ll(n), P(u), P(tail), P(bsic), P(ra))
        return u, tail, bsic, ra

    def analyze_52m(self, burst, tsc, thr=3.0, max_toa=3, request=True):
        burst = np.ascontiguousarray(burst, np.complex64)
        amp = np.zeros(1, np.complex64); toa = np.zeros(1, np.float32)
        chan = np.zeros(6 * self.sps, np.complex64); off = np.zeros(1, np.float32)
        ok = self.lib.emu_analyze_52m(P(burst), c_i(burst.size), c_i(tsc), c_f(thr), ctypes.c_uint(max_toa), c_i(int(request)),
                                      P(amp), P(toa), P(chan), P(off))
        return bool(ok), amp[0], toa[0], chan, off[0]

    def analyze_52m_tiled(self, burst, tsc, thr=3.0, max_toa=3, request=True):
        """the same through k_detect_52m's tile layout (only the search window staged, everything else poisoned); sps 1"""
        burst = np.ascontiguousarray(burst, np.complex64)
        amp = np.zeros(1, np.complex64); toa = np.zeros(1, np.float32)
        chan = np.zeros(6, np.complex64); off = np.zeros(1, np.float32)
        ok = self.lib.emu_analyze_52m_tiled(P(burst), c_i(burst.size), c_i(tsc), c_f(thr), ctypes.c_uint(max_toa), c_i(int(request)),
                                            P(amp), P(toa), P(chan), P(off))
        return bool(ok), amp[0], toa[0], chan, off[0]

    def energy_detect_52m(self, v, win, thr):
        v = np.ascontiguousarray(v, np.complex64)
        avg = np.zeros(1, np.float32)
        ok = self.lib.emu_energy_detect_52m(P(v), c_i(v.size), ctypes.c_uint(win), c_f(thr), P(avg))
        return bool(ok), avg[0]

    def rx_rach_batch(self, bursts, lens, detect_thr=5.0, tiles=True):
        bursts = np.ascontiguousarray(bursts, np.complex64)
        n, pitch = bursts.shape
        lens = np.ascontiguousarray(lens, np.int32)
        r = dict(flag=np.zeros(n, np.int32), amp=np.zeros(n, np.complex64), toa=np.zeros(n, np.float32),
                 soft=np.zeros((n, 160 * (1 if tiles else self.sps)), np.float32))
        self.lib.emu_rach(P(bursts), c_ll(pitch), P(lens), c_ll(0), c_ll(n), c_f(detect_thr), c_i(self.sps),
                          c_i(int(tiles)), P(r["flag"]), P(r["amp"]), P(r["toa"]), P(r["soft"]), c_i(r["soft"].shape[1]))
        return r

    def analyze_batch(self, bursts, lens, tsc, detect_thr=3.0, request=True):
        bursts = np.ascontiguousarray(bursts, np.complex64)
        n, pitch = bursts.shape
        lens = np.ascontiguousarray(lens, np.int32)
        tsc = np.ascontiguousarray(tsc, np.uint8)
        r = dict(flag=np.zeros(n, np.int32), amp=np.zeros(n, np.complex64), toa=np.zeros(n, np.float32),
                 chan=np.zeros((n, 6 * self.sps), np.complex64), off=np.zeros(n, np.float32))
        self.lib.emu_analyze(P(bursts), c_ll(pitch), P(lens), c_ll(0), P(tsc), c_ll(n), c_f(detect_thr), c_i(int(request)),
                             c_i(self.sps), P(r["flag"]), P(r["amp"]), P(r["toa"]), P(r["chan"]), P(r["off"]))
        return r

    def peak_detect(self, v, grid=False):
        v = np.ascontiguousarray(v, np.complex64)
        pk = np.zeros(1, np.complex64)
        idx, avg = c_f(), c_f()
        (self.lib.emu_peak_detect_grid if grid else self.lib.emu_peak_detect)(P(v), c_i(v.size), P(pk), ctypes.byref(idx),
                                                                              ctypes.byref(avg))
        return pk[0], idx.value, avg.value

    def delay_vector(self, v, delay):
        v = np.ascontiguousarray(v, np.complex64).copy()
        self.lib.emu_delay_vector(P(v), c_i(v.size), c_f(delay))
        return v

    def design_dfe(self, chan, snr, nf=7, fixed=False):
        chan = np.ascontiguousarray(chan, np.complex64)
        w, b = np.zeros(nf, np.complex64), np.zeros(max(chan.size - 1, 1), np.complex64)
        self.lib.emu_design_dfe(P(chan), c_i(chan.size), c_f(snr), c_i(nf), P(w), P(b), c_i(int(fixed)))
        return w, b[:chan.size - 1]

    def equalize(self, burst, toa, w, b):
        burst = np.ascontiguousarray(burst, np.complex64).copy()
        w, b = np.ascontiguousarray(w, np.complex64), np.ascontiguousarray(b, np.complex64)
        soft = np.zeros(burst.size, np.float32)
        self.lib.emu_equalize(P(burst), c_i(burst.size), c_f(toa), P(w), c_i(w.size), P(b), c_i(b.size), P(soft))
        return soft, burst

    def modulate(self, bits, guard):
        bits = np.ascontiguousarray(bits, np.uint8)
        out = np.zeros(self.sps * (bits.size + guard), np.complex64)
        self.lib.emu_modulate(P(bits), c_i(bits.size), c_i(guard), P(out))
        return out

    def rx_resample_stream(self, raw):
        raw = np.ascontiguousarray(raw, np.complex64)
        nch = raw.size // 864
        out = np.zeros(nch * 585, np.complex64)
        self.lib.emu_resample_rx(P(raw), c_i(0), c_ll(nch), P(out))
        return out

    def rx_resample_stream_v2(self, raw, has_history=False, offset=0):
        """the tuned kernel's logic; with has_history the 192 samples before raw[offset] are read as history"""
        raw = np.ascontiguousarray(raw, np.complex64)
        nch = (raw.size - offset) // 864
        out = np.zeros(nch * 585, np.complex64)
        self.lib.emu_resample_rx_v2(ctypes.c_void_p(raw.ctypes.data + 8 * offset), c_i(int(has_history)), c_ll(nch), P(out))
        return out

    def tx_resample_stream(self, x):
        x = np.ascontiguousarray(x, np.complex64)
        nch = x.size // 585
        out = np.zeros((nch * 864, 2), np.int16)
        self.lib.emu_resample_tx(P(x), c_i(0), c_ll(nch), P(out))
        return out
