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
