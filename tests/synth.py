"""Seeded synthetic GSM bursts for the parity tests (test infrastructure; numpy only).

Follows SURVEY.md 8(d): normal bursts = 3 tail + 58 data + 26 TSC + 58 data + 3 tail (+guard), access
bursts = 8 ext-tail + 41 sync + 36 data + 3 tail (+guard); a burst is modulated by the caller-supplied
modulator (oracle or product), then amplitude/phase, fractional delay, an optional 2-tap channel and
AWGN are applied in float64 and rounded once to complex64 -- the same array is then fed to every
implementation, so how it was made does not matter for parity.
"""
import numpy as np

TSC = ["00100101110000100010010111", "00101101110111100010110111", "01000011101110100100001110",
       "01000111101101000100011110", "00011010111001000001101011", "01001110101100000100111010",
       "10100111110110001010011111", "11101111000100101110111100"]   # GSM 05.02 5.2.3
RACH_SYNC = "01001011011111111001100110101010001111000"              # GSM 05.02 5.2.7
RACH_EXT_TAIL = "00111010"


def bits_of(s):
    return np.frombuffer(s.encode(), np.uint8) - ord("0")


def normal_burst_bits(rng, tsc):
    """148 bits: 000 | 58 | TSC(26) | 58 | 000 ; stealing flags are just data here."""
    b = np.zeros(148, np.uint8)
    b[3:61] = rng.integers(0, 2, 58)
    b[61:87] = bits_of(TSC[tsc])
    b[87:145] = rng.integers(0, 2, 58)
    return b


def access_burst_bits(rng):
    """88 bits used of 148: ext tail(8) | sync(41) | data(36) | 000 ; rest of the slot is guard (zeros bits -> we
    leave them unmodulated by passing only 88 bits and a longer guard)."""
    b = np.zeros(88, np.uint8)
    b[0:8] = bits_of(RACH_EXT_TAIL)
    b[8:49] = bits_of(RACH_SYNC)
    b[49:85] = rng.integers(0, 2, 36)
    return b


def frac_delay(x, d):
    """band-limited delay by d samples (float64 windowed-sinc, 41 taps); only used to MAKE inputs"""
    io = int(np.floor(d))
    f = d - io
    n = np.arange(-20, 21)
    h = np.sinc(n - f) * np.hamming(41)
    y = np.convolve(x.astype(np.complex128), h)[20:20 + x.size]
    out = np.zeros_like(y)
    if io >= 0:
        out[io:] = y[:x.size - io]
    else:
        out[:io] = y[-io:]
    return out


def impair(rng, x, amp=1000.0, phase=None, delay=0.0, chan2=None, snr_db=None):
    y = x.astype(np.complex128)
    if chan2 is not None:
        y = y + chan2 * np.concatenate([[0], y[:-1]])
    if delay != 0.0:
        y = frac_delay(y, delay)
    if phase is None:
        phase = rng.uniform(0, 2 * np.pi)
    y = y * amp * np.exp(1j * phase)
    if snr_db is not None:
        sigma = amp / np.sqrt(10 ** (snr_db / 10.0))
        y = y + sigma / np.sqrt(2) * (rng.standard_normal(y.size) + 1j * rng.standard_normal(y.size))
    return y.astype(np.complex64)


def make_normal_batch(modulate, n, seed=1, mixed_tsc=True, pitch=160, snr=(10, 30), amp=(500, 8000),
                      max_delay=3.0, multipath=0.5, noise_only=0.0):
    """n normal bursts, SURVEY config-4 style.  Returns (bursts[n,pitch] c64, lens[n] i32, tsc[n] u8, bits[n,148])."""
    rng = np.random.default_rng(seed)
    bursts = np.zeros((n, pitch), np.complex64)
    lens = np.zeros(n, np.int32)
    tscs = np.zeros(n, np.uint8)
    bits = np.zeros((n, 148), np.uint8)
    for i in range(n):
        tn = i % 8
        arfcn = i // 8
        tsc = (arfcn % 8) if mixed_tsc else 0
        guard = 8 + (tn % 4 == 0)
        b = normal_burst_bits(rng, tsc)
        x = modulate(b, guard)
        a = float(np.exp(rng.uniform(np.log(amp[0]), np.log(amp[1]))))
        ch = 0.4 * np.exp(1j * rng.uniform(0, 2 * np.pi)) if rng.random() < multipath else None
        y = impair(rng, x, amp=a, delay=rng.uniform(0, max_delay), chan2=ch, snr_db=rng.uniform(*snr))
        if rng.random() < noise_only:
            y = (a * 0.3 * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))).astype(np.complex64)
        bursts[i, :x.size] = y
        lens[i] = x.size
        tscs[i] = tsc
        bits[i] = b
    return bursts, lens, tscs, bits


def make_rach_batch(modulate, n, seed=2, pitch=160, snr=(-5, 20), max_delay=63):
    rng = np.random.default_rng(seed)
    bursts = np.zeros((n, pitch), np.complex64)
    lens = np.zeros(n, np.int32)
    bits = np.zeros((n, 88), np.uint8)
    delays = np.zeros(n)
    for i in range(n):
        total = 157 if (i % 4 == 0) else 156
        b = access_burst_bits(rng)
        x = modulate(b, total - 88)
        d = rng.integers(0, max_delay + 1) + rng.uniform(0, 1)
        y = impair(rng, x, amp=1000.0, delay=d, snr_db=rng.uniform(*snr))
        bursts[i, :total] = y
        lens[i] = total
        bits[i] = b
        delays[i] = d
    return bursts, lens, bits, delays


def make_edge_batch(o, n=70, seed=9):
    """hand-made hard cases for the normal-burst path: flat input, pure noise, impulses in / at the edge of the
    correlation window, clean bursts with zero / early / late timing, ragged count (not a multiple of 32).
    `o` supplies modulate() and delay_vector() (the oracle)."""
    rng = np.random.default_rng(seed)
    bursts = np.zeros((n, 160), np.complex64)
    lens = np.where(np.arange(n) % 4 == 0, 157, 156).astype(np.int32)
    tsc = (np.arange(n) % 8).astype(np.uint8)
    bursts[1] = 1e-3
    bursts[2, :156] = (rng.standard_normal(156) + 1j * rng.standard_normal(156)) * 1e4
    bursts[3, 60] = 5000
    bursts[4, 91] = 5000 + 1j
    for i, (t, d) in enumerate([(3, 0.0), (3, -2.5), (3, 9.25), (0, -7.0), (0, 7.5), (2, 12.0), (6, -11.3), (7, 3.0),
                                (0, 5 / 512), (0, 6 / 512), (2, -4 + 3 / 512), (6, 1.999), (7, -0.001)]):
        b = normal_burst_bits(rng, t)
        k = 5 + i
        tsc[k] = t
        x = o.modulate(b, 8 + (k % 4 == 0)) * np.complex64(3000 * np.exp(1j * i))
        bursts[k, :lens[k]] = o.delay_vector(x, d)
    for i in range(18, n):
        if i % 3 == 0:
            b = normal_burst_bits(rng, tsc[i])
            x = o.modulate(b, 8 + (i % 4 == 0)) * np.complex64(rng.uniform(30, 30000))
            y = o.delay_vector(x, rng.uniform(-6, 6))
            y = y + (rng.standard_normal(y.size) + 1j * rng.standard_normal(y.size)) * rng.uniform(0, 300)
            bursts[i, :lens[i]] = y.astype(np.complex64)
        else:
            bursts[i, :lens[i]] = (rng.standard_normal(lens[i]) + 1j * rng.standard_normal(lens[i])) * rng.uniform(1, 3000)
    return bursts, lens, tsc


def make_trx_batch(modulate, corr_type, nframes, tsc, chan_type, fn0=0, seed=5, pitch=160):
    """Received slots for the caller-policy tests, laid out [frame][arfcn][tn]: on TSC slots mostly normal bursts of
    the ARFCN's midamble (amplitudes straddling the 250 energy threshold, a per-timeslot 2-tap channel), sometimes
    plain noise (false detections) or near-silence; on RACH slots access bursts or noise; junk elsewhere.
    corr_type(chan_type, fn) -> 0 off / 1 TSC / 2 RACH / 3 idle.  Returns bursts[(nframes*narfcn*8), pitch] c64."""
    rng = np.random.default_rng(seed)
    tsc = np.asarray(tsc)
    A = tsc.size
    chan_type = np.asarray(chan_type).reshape(A, 8)
    out = np.zeros((nframes * A * 8, pitch), np.complex64)
    chan = 0.4 * np.exp(2j * np.pi * rng.random((A, 8))) * (rng.random((A, 8)) < 0.5)
    for f in range(nframes):
        for a in range(A):
            for tn in range(8):
                i = (f * A + a) * 8 + tn
                n = 157 if tn % 4 == 0 else 156
                c = corr_type(int(chan_type[a, tn]), fn0 + f)
                u = rng.random()
                noise = lambda p: (p * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
                if c == 1 and u < 0.7:
                    x = modulate(normal_burst_bits(rng, int(tsc[a])), n - 148)
                    amp = float(np.exp(rng.uniform(np.log(60), np.log(5000))))
                    ch = chan[a, tn] if chan[a, tn] != 0 else None
                    out[i, :n] = impair(rng, x, amp=amp, delay=rng.uniform(0, 2), chan2=ch, snr_db=rng.uniform(8, 30))
                elif c == 2 and u < 0.5:
                    x = modulate(access_burst_bits(rng), n - 88)
                    amp = float(np.exp(rng.uniform(np.log(100), np.log(3000))))
                    out[i, :n] = impair(rng, x, amp=amp, delay=rng.integers(0, 40) + rng.random(), snr_db=rng.uniform(0, 25))
                elif u < 0.85:
                    out[i, :n] = noise(float(np.exp(rng.uniform(np.log(30), np.log(2000)))))
                else:
                    out[i, :n] = noise(1.0)
    return out
