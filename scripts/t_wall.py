"""Where the wall time of a bench step goes: reset / push_all / view_frames / release, per call."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from usrp_nfc_b200 import _cabi
codes, lens, params = bench.build_schedule(bench.RATE, 2024)
n = int(1e10); n -= n % (13560 * 4)
x = torch.empty(n, dtype=torch.float32, device="cuda")
chan = dict(carrier=0.5, pause=0.015, tag_high=1.07, noise=0.003, fade=0.05, fade_period=round(bench.RATE * 0.02))
_cabi.synth_render(x, codes, lens, seed=99, as_envelope=True, device=0, first_index=0, **chan)
torch.cuda.synchronize()
s = _cabi.Stream(bench.RATE, hi_val=bench.HI_VAL, outputs=_cabi.OUT_FRAMES, device=0, **params)
s.set_tuning(slab_len=1 << 30)
for it in range(5):
    t0 = time.perf_counter(); s.reset()
    t1 = time.perf_counter(); s.push_all(x)
    t2 = time.perf_counter(); fr, b0, b1 = s.view_frames()
    t3 = time.perf_counter(); s.release_frames()
    t4 = time.perf_counter()
    st = s.stats()
    print("reset %.2f push %.2f view %.2f release %.2f total %.2f ms | kernel_ms %.2f" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3, st["kernel_ms"]))
    s.reset_stats()
