"""Where a capture's time goes in the batch path: per-call wall times inside the workers (diagnostic)."""
import sys, time, threading
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from usrp_nfc_b200 import _cabi

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
NCAP = 256
rate = 13.56e6
codes, lens, p = bench.build_schedule(rate, 2024)
pool = []
for u in range(8):
    x = torch.empty(4_000_000, dtype=torch.float32, device="cuda")
    _cabi.synth_render(x, codes, lens, carrier=0.5, pause=0.015, tag_high=1.07, noise=0.003, fade=0.05, fade_period=round(rate * 0.02),
                       seed=500 + u, as_envelope=True, device=0, first_index=u * 7919 * 4096)
    pool.append(x)
torch.cuda.synchronize()
acc = np.zeros(5)
lock = threading.Lock()
nxt = [0]

def work():
    s = _cabi.Stream(rate, hi_val=1.09, outputs=_cabi.OUT_FRAMES, **p)
    s.set_tuning(slab_len=1 << 28)
    s.set_wait_mode(True)
    loc = np.zeros(5)
    while True:
        with lock:
            i = nxt[0]; nxt[0] += 1
        if i >= NCAP:
            break
        x = pool[i % 8]
        t0 = time.perf_counter(); s.reset(); s.set_thresholds(0.1, 1.05 + 0.01 * (i % 6))
        t1 = time.perf_counter(); u, _ = s.push(x)
        t2 = time.perf_counter(); s.push(x[u:])
        t3 = time.perf_counter(); s.drain_frames_flat()
        t4 = time.perf_counter()
        if i >= W:
            loc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3, 1]
    with lock:
        acc[:] += loc
    s.close()

t0 = time.perf_counter()
ths = [threading.Thread(target=work) for _ in range(W)]
[t.start() for t in ths]; [t.join() for t in ths]
wall = time.perf_counter() - t0
n = acc[4]
print("workers %d: wall %.1f ms for %d captures (%.3f ms each); per capture inside a worker: reset+thresholds %.3f, warm-up push %.3f, push %.3f, drain %.3f ms"
      % (W, wall * 1e3, NCAP, wall * 1e3 / NCAP, *(acc[:4] / n * 1e3)))
