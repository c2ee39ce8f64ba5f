import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from usrp_nfc_b200 import _cabi
codes, lens, params = bench.build_schedule(13.56e6, 2024)
n = int(4e9); L = params["av_window"]
x = torch.empty(n, dtype=torch.float32, device="cuda")
chan = dict(carrier=0.5, pause=0.015, tag_high=1.07, noise=0.003, fade=0.05, fade_period=round(13.56e6 * 0.02))
_cabi.synth_render(x, codes, lens, seed=99, as_envelope=True, device=0, first_index=0, **chan)
torch.cuda.synchronize()
s = _cabi.Stream(13.56e6, hi_val=1.09, outputs=_cabi.OUT_FRAMES, device=0, **params)
for i in range(3):
    s.reset()
    t0 = time.perf_counter(); s.push_all(x); t1 = time.perf_counter(); fr, bits = s.drain_frames_flat(); t2 = time.perf_counter()
    print("push %.2f ms drain %.2f ms frames %d bits %d" % ((t1-t0)*1e3, (t2-t1)*1e3, len(fr), len(bits)))
