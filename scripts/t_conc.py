import sys, time, threading, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from usrp_nfc_b200 import _cabi
rate = 13.56e6
codes, lens, params = bench.build_schedule(rate, 2024)
ns = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
chan = dict(carrier=0.5, pause=0.015, tag_high=1.07, noise=0.003, fade=0.05, fade_period=round(rate * 0.02))
xs = []
for u in range(8):
    x = torch.empty(ns, dtype=torch.float32, device="cuda")
    _cabi.synth_render(x, codes, lens, seed=500 + u, as_envelope=True, device=0, first_index=u * 7919 * 4096, **chan)
    xs.append(x)
torch.cuda.synchronize()
def run(nthreads, reps):
    streams = [_cabi.Stream(rate, hi_val=1.09, outputs=_cabi.OUT_FRAMES, **params) for _ in range(nthreads)]
    times = [[] for _ in range(nthreads)]
    def work(k):
        s = streams[k]
        for r in range(reps):
            t0 = time.perf_counter()
            s.reset(); s.push_all(xs[(k + r) % 8]); s.drain_frames_flat(reuse=True)
            times[k].append(time.perf_counter() - t0)
    for k in range(nthreads): work_k = None
    # warm
    for k in range(nthreads):
        s = streams[k]; s.reset(); s.push_all(xs[k % 8]); s.drain_frames_flat(reuse=True)
    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(k,)) for k in range(nthreads)]
    [t.start() for t in ths]; [t.join() for t in ths]
    wall = time.perf_counter() - t0
    per = np.mean([np.mean(t) for t in times]) * 1e3
    print("threads %d: wall %.1f ms for %d decodes (%.2f ms each amortised), per-call latency %.2f ms" % (nthreads, wall * 1e3, nthreads * reps, wall * 1e3 / (nthreads * reps), per))
    for s in streams: s.close()
for nt in (1, 2, 4, 8):
    run(nt, 16)
