"""Where the time of a work()-sized call goes: 8192-item pushes of 16-bit PCM at 2 MS/s, each followed by the two calls a
drain makes; cumulative time per ABI call.  python scripts/calls_probe.py [calls]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from usrp_nfc_b200 import _cabi, synth  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rate = 2e6
frames = synth.load_sessions()["ultralight"]
pcm = synth.capture(frames, rate, 3, sessions=8)
pcm = np.ascontiguousarray(np.tile(pcm, 1 + calls * 8192 // len(pcm))[:calls * 8192])
s = _cabi.Stream(rate, hi_val=1.09, outputs=_cabi.OUT_FRAMES, input_kind=_cabi.IN_PCM_S16)
L = _cabi.lib()
t = {"push": 0.0, "count": 0.0, "pending": 0.0, "drain": 0.0}
nfr = 0
for rep in range(2):
    for k in t:
        t[k] = 0.0
    s.reset()
    for i in range(calls):
        x = pcm[i * 8192:(i + 1) * 8192]
        a = time.perf_counter()
        s.push(x)
        b = time.perf_counter()
        n = L.nfc_stream_drain_frames(s._h, None, 0, None, 0)
        c = time.perf_counter()
        L.nfc_stream_pending_frame_bits(s._h)
        d = time.perf_counter()
        if n:
            nfr += len(s.drain_frames_flat()[0])
        e = time.perf_counter()
        t["push"] += b - a
        t["count"] += c - b
        t["pending"] += d - c
        t["drain"] += e - d
    print(os.environ.get("USRP_NFC_B200_LIB", "in-tree"), rep, {k: "%.3f ms/call" % (v * 1e3 / calls) for k, v in t.items()}, "frames", nfr)
