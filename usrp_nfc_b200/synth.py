"""Synthetic ISO 14443A reader+tag traffic on the host (numpy) -- SURVEY.md section 8(d).

Frames are line-coded with the same pulse grammar as the reference's TX side
(miller.py:200-233 `miller_encoder`, manchester.py:64-79 `manchester_encoder`,
pulse constants utilities.py:16-23) and rendered the way binary_src.work does
(binary_src.py:64-103): one amplitude level per pulse, held for its duration.
The envelope carries no 847 kHz subcarrier: a modulated Manchester half-bit is one
contiguous above-carrier run, which is what the reference's slicer requires
(SURVEY.md section 0, fact 4).

This is the input generator for tests and bench.py; the device-side generator in
csrc/synth.cu renders the same pulse schedule for captures too large for the host.
"""
import json
import os

import numpy as np

# utilities.py:16-23
FULL = 9.44
ZERO = 3.00
HALF = FULL / 2
ZERO_REM = FULL - ZERO
ONE_REM = HALF - ZERO
ONE_HALF = FULL + HALF

TAG_TO_READER = 0  # packets.py:19-20
READER_TO_TAG = 1

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def miller_encode(bits):
    """Reader->tag pulses [(level, dur_us)], start and end bits added (miller.py:206-233)."""
    one = [(1, HALF), (0, ZERO), (1, ONE_REM)]
    zero0 = [(0, ZERO), (1, ZERO_REM)]
    zero1 = [(1, FULL)]
    durs = list(zero0)
    last_bit = 0
    for bit in list(bits) + [0]:
        cur = one
        if bit == 0:
            cur = zero0 if last_bit == 0 else zero1
        last_bit = bit
        last_pulse, last_dur = durs[-1]
        if cur[0][0] == last_pulse:
            durs[-1] = (last_pulse, cur[0][1] + last_dur)
            durs.extend(cur[1:])
        else:
            durs.extend(cur)
    return durs


def manchester_encode(bits):
    """Tag->reader pulses [(level, dur_us)], start bit added (manchester.py:66-79)."""
    durs = [(1, HALF), (0, HALF)]
    last = 0
    for bit in bits:
        if bit == last:
            durs[-1] = (bit, FULL)
            last = 1 - last
            durs.append((last, HALF))
        else:
            durs.append((1 - last, HALF))
            durs.append((last, HALF))
    return durs


def bytes_to_bits(data, parity=True):
    """LSB-first bits with odd parity after each byte (utilities.py:52-63)."""
    out = []
    for b in data:
        ones = 0
        for i in range(8):
            bit = (b >> i) & 1
            ones += bit
            out.append(bit)
        if parity:
            out.append(1 - (ones & 1))
    return out


def load_sessions():
    """Session scripts: on-air frames [(packet_type, [bits])] of the reference's two logged sessions
    (outputs/ultralight.out, outputs/1k_with_enc.out), extracted by oracle/gen_golden.py."""
    with open(os.path.join(_DATA, "sessions.json")) as f:
        raw = json.load(f)
    return {k: [(int(t), [int(c) for c in s]) for t, s in v] for k, v in raw.items()}


class Channel(object):
    """Amplitude model of SURVEY.md 8(d): carrier 0.5 FS, pause depth, tag-high ratio, noise, fade."""

    def __init__(self, carrier=0.5, pause=0.02, tag_high=1.08, noise=0.003, fade=0.0, fade_period_us=20000.0):
        self.carrier, self.pause, self.tag_high = carrier, pause, tag_high
        self.noise, self.fade, self.fade_period_us = noise, fade, fade_period_us


def schedule(frames, samp_rate, rng, lead_us=None, reader_gap_us=86.0, tag_gap_us=(200.0, 600.0),
             tail_us=1000.0, av_window=2000):
    """Pulse schedule for one session: arrays (level_code, n_samples) with cumulative rounding.

    level_code: 0 carrier, 1 reader pause, 2 tag high.
    """
    us = 1e6 / samp_rate
    if lead_us is None:
        lead_us = (av_window + 200) * us + 300.0
    codes, ends = [], []
    t = 0.0

    def put(code, dur):
        nonlocal t
        t += dur
        codes.append(code)
        ends.append(t)

    put(0, lead_us)
    for ptype, bits in frames:
        if ptype == READER_TO_TAG:
            for level, dur in miller_encode(bits):
                put(0 if level else 1, dur)
            put(0, reader_gap_us)
        else:
            for level, dur in manchester_encode(bits):
                put(2 if level else 0, dur)
            put(0, float(rng.uniform(*tag_gap_us)))
    put(0, tail_us)
    edges = np.rint(np.asarray(ends) * (samp_rate / 1e6)).astype(np.int64)
    lens = np.diff(np.concatenate(([0], edges)))
    codes = np.asarray(codes, dtype=np.int8)
    keep = lens > 0
    return codes[keep], lens[keep]


def render(codes, lens, samp_rate, rng, channel=None):
    """Pulse schedule -> int16 PCM (what a 16-bit WAV recording of the magnitude would hold)."""
    ch = channel or Channel()
    mult = np.array([1.0, ch.pause / ch.carrier, ch.tag_high])[codes]
    amp = np.repeat(mult, lens) * ch.carrier
    n = amp.size
    if ch.noise:
        amp *= 1.0 + ch.noise * rng.standard_normal(n)
    if ch.fade:
        tt = np.arange(n) * (1e6 / samp_rate)
        amp *= 1.0 + ch.fade * np.sin(2 * np.pi * tt / ch.fade_period_us)
    return np.clip(np.rint(amp * 32767.0), -32768, 32767).astype(np.int16)


def pcm_to_float(pcm):
    """What blocks.wavfile_source emits for 16-bit PCM (decoder.py:25): sample / 0x7FFF in float32.
    (GNU Radio is not under the reference checkout; the constant is restated from GR 3.7 and unpinned.)"""
    return pcm.astype(np.float32) / np.float32(32767.0)


def envelope(x):
    """float_to_complex(1) + complex_to_mag_squared(1) with im = 0 (decoder.py:26-28): RN32(x*x)."""
    x = np.asarray(x, dtype=np.float32)
    return x * x


def capture(frames, samp_rate, seed, channel=None, av_window=2000, sessions=1, idle_us=(1000.0, 5000.0), **kw):
    """int16 PCM of `sessions` repetitions of a session script with U[idle] gaps between them."""
    rng = np.random.default_rng(seed)
    parts = []
    for s in range(sessions):
        lead = None if s == 0 else float(rng.uniform(*idle_us))
        c, l = schedule(frames, samp_rate, rng, lead_us=lead, av_window=av_window, **kw)
        parts.append(render(c, l, samp_rate, rng, channel))
    return np.concatenate(parts)


def rate_params(samp_rate):
    """Constructor parameters per sample rate: max_len = 25 us, av_window = 1 ms."""
    if abs(samp_rate - 2e6) < 1:
        return dict(av_window=2000, max_len=50)
    return dict(av_window=int(round(samp_rate * 1e-3)), max_len=int(round(samp_rate * 25e-6)))


def random_frames(rng, n_frames, max_bytes=18):
    """Alternating reader/tag frames with random payloads (for seeded differential tests)."""
    out = []
    for i in range(n_frames):
        nb = int(rng.integers(1, max_bytes + 1))
        data = rng.integers(0, 256, nb).tolist()
        bits = bytes_to_bits(data)
        if rng.random() < 0.1:
            bits = bits[:7]
        out.append((READER_TO_TAG if i % 2 == 0 else TAG_TO_READER, bits))
    return out
