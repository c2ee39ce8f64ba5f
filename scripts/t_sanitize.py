"""Small decode through the streaming kernel for compute-sanitizer (racecheck / memcheck)."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from usrp_nfc_b200 import _cabi, synth
rate = 13.56e6
p = synth.rate_params(rate)
frames = synth.load_sessions()["ultralight"]
pcm = synth.capture(frames, rate, 5, channel=synth.Channel(pause=0.015, tag_high=1.07, fade=0.05), av_window=p["av_window"], sessions=2)
x = synth.envelope(synth.pcm_to_float(pcm))
print("samples", x.size)
for kind, data in ((_cabi.IN_ENVELOPE_F32, x), (_cabi.IN_PCM_S16, pcm)):
    s = _cabi.Stream(rate, hi_val=1.09, outputs=_cabi.OUT_ALL, input_kind=kind, **p)
    s.set_tuning(seg_len=100000, halo=2 * 13560, slab_len=1 << 28)  # several speculative segments, seams, repairs
    s.push_all(data)
    ev = s.drain_events(); fr, bits = s.drain_frames()
    st = s.stats()
    print("kind", kind, "events", len(ev), "frames", len(fr), {k: st[k] for k in ("segments", "seam_mismatches", "fast_tiles", "exact_tiles", "repeated_passes", "fixpoint_tiles")})
    s.close()

# a stream that provokes the rare paths too: samples hovering around the HIGH threshold (fix-point pass), spikes right
# after pauses (hysteresis -> exact path), level steps (coarser fixed-point step)
rng = np.random.default_rng(11)
L, mx, n = 8192, 50, 260000
y = (0.25 * (1 + 0.01 * rng.standard_normal(n))).astype(np.float32)
i = L + 10
while i < n - 4 * mx - 10:
    kind = rng.integers(0, 5)
    ln = min(int(rng.choice([1, 2, mx, 2 * mx + 1, 6, 700])), n - i - 1)
    if kind == 0:
        y[i:i + ln] = 1e-4
        i += ln
        if rng.random() < 0.7:
            k = int(rng.integers(0, mx + 4))
            y[i + k: i + k + 2] = 0.4
    elif kind == 2:
        y[i:i + ln] = 0.2725 * (1 + 0.002 * rng.standard_normal(ln))
        i += ln
    elif kind == 3 and rng.random() < 0.05:
        y[i:] *= np.float32(rng.choice([0.7, 1.4]))
    i += int(rng.integers(1, 6 * mx))
s = _cabi.Stream(13.56e6, hi_val=1.09, outputs=_cabi.OUT_ALL, av_window=L, max_len=mx)
s.set_tuning(seg_len=65536, halo=4 * L)
s.push_all(y)
ev = s.drain_events(); fr, bits = s.drain_frames()
st = s.stats()
print("corner events", len(ev), "frames", len(fr), {k: st[k] for k in ("segments", "seam_mismatches", "fast_tiles", "exact_tiles", "repeated_passes", "fixpoint_tiles")})
s.close()

# the small-window variant of the streaming kernel (tiles of 512 samples): the reference's defaults at 2 MS/s, several
# slabs (queued chains), and a corner stream with its hysteresis cases
rate = 2e6
frames = synth.load_sessions()["classic1k"]
pcm = synth.capture(frames, rate, 7, channel=synth.Channel(pause=0.015, tag_high=1.07, fade=0.05), av_window=2000, sessions=2)
x = synth.envelope(synth.pcm_to_float(pcm))
s = _cabi.Stream(rate, hi_val=1.09, outputs=_cabi.OUT_ALL)
s.set_tuning(seg_len=40000, halo=8000, slab_len=1 << 18)
s.push_all(x)
ev = s.drain_events(); fr, bits = s.drain_frames()
st = s.stats()
print("small kind 0 events", len(ev), "frames", len(fr), {k: st[k] for k in ("segments", "seam_mismatches", "fast_tiles", "exact_tiles", "repeated_passes", "fixpoint_tiles", "pipe_tiles", "overflow_retries")})
s.close()
L, n = 2000, 120000
y = (0.25 * (1 + 0.01 * rng.standard_normal(n))).astype(np.float32)
i = L + 10
while i < n - 4 * mx - 10:
    kind = rng.integers(0, 5)
    ln = min(int(rng.choice([1, 2, mx, 2 * mx + 1, 6, 700])), n - i - 1)
    if kind == 0:
        y[i:i + ln] = 1e-4
        i += ln
        if rng.random() < 0.7:
            k = int(rng.integers(0, mx + 4))
            y[i + k: i + k + 2] = 0.4
    elif kind == 2:
        y[i:i + ln] = 0.2725 * (1 + 0.002 * rng.standard_normal(ln))
        i += ln
    i += int(rng.integers(1, 6 * mx))
s = _cabi.Stream(2e6, hi_val=1.09, outputs=_cabi.OUT_ALL, av_window=L, max_len=mx)
s.set_tuning(seg_len=16384, halo=4 * L)
s.push_all(y)
ev = s.drain_events(); fr, bits = s.drain_frames()
st = s.stats()
print("small corner events", len(ev), "frames", len(fr), {k: st[k] for k in ("segments", "seam_mismatches", "fast_tiles", "exact_tiles", "repeated_passes", "fixpoint_tiles")})
s.close()
