"""Time-shard stitching with the real CUDA engine: 2 and 3 ranks (gloo rendezvous, all on cuda:0).
The merged frame list must equal the oracle's decode of the whole capture."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, halo_windows, rate, q, view=False):
    import torch.distributed as dist
    from oracle import oracle
    from usrp_nfc_b200 import _cabi, sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = synth.rate_params(rate)
    frames = synth.load_sessions()["classic1k"]
    pcm = synth.capture(frames, rate, 31, channel=synth.Channel(pause=0.02, tag_high=1.07, fade=0.05),
                        av_window=p["av_window"], sessions=2)
    x = synth.envelope(synth.pcm_to_float(pcm))
    eng = _cabi.Stream(rate, hi_val=1.09, outputs=_cabi.OUT_FRAMES, **p)
    eng.set_tuning(seg_len=16 * p["av_window"], halo=4 * p["av_window"])
    # the frame offsets alone: in shared memory of the node when the view form is used (what bench.py does), else over the
    # process group
    shared = sharding.SharedFrameIndex(eng, 4096, dist) if view else None
    res = sharding.decode_time_sharded(eng, lambda a, b: x[a:b], x.size, p["av_window"], _cabi.State, dist=dist,
                                       halo_windows=halo_windows, flat="view" if view else False)
    if view:  # zero-copy bulk form (what bench.py uses): records + the engine's own bit buffers, released by the caller
        b0, b1 = res["bits"]
        mine = []
        for r in res["records"]:
            buf = b0 if int(r["type"]) == 0 else b1
            mine.append((int(r["pos"]) + res["pos_offset"], int(r["type"]),
                         buf[int(r["bit_off"]): int(r["bit_off"]) + int(r["nbits"])].copy()))
        assert len(mine) == res["n_frames"]
        index = sharding.gather_frame_records(eng, res["pos_offset"], dist, shared=shared)
        idx = index.unpack() if index is not None else None  # (a copy: the segments are unmapped below)
        assert shared.ok
        eng.release_frames()
        shared.close()
    else:
        mine = res["frames"]
        rec = np.zeros(len(mine), dtype=_cabi.FRAME_DTYPE)
        for i, (pp, tt, bb) in enumerate(mine):
            rec[i] = (pp, 0, len(bb), tt)
        index = sharding.gather_frame_records(rec, 0, dist)
        idx = index.unpack() if index is not None else None
    merged = sharding.gather_frames(mine, dist)
    if rank == 0:
        want = oracle.decode_capture(x, rate, hi_val=1.09, **p)
        ok = len(merged) == len(want["frames"])
        ok = ok and len(idx) == len(want["frames"]) and np.array_equal(idx["pos"], want["frames"]["pos"])
        ok = ok and np.array_equal(idx["nbits"], want["frames"]["nbits"]) and np.array_equal(idx["type"], want["frames"]["type"])
        ok = ok and all(pp == int(w["pos"]) and t == int(w["type"]) and np.array_equal(b, wb)
                        for (pp, t, b), w, wb in zip(merged, want["frames"], want["frame_bits"]))
        q.put((ok, len(merged), len(want["frames"])))
    q.put(("rank", rank, res["repaired"], res["seam_ok"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,halo_windows,rate,view", [(2, 16, 2e6, False), (3, 1, 2e6, False), (2, 2, 13.56e6, False),
                                                          (3, 2, 13.56e6, True), (2, 2, 20e6, True)])
def test_gpu_time_shards_stitch_exactly(world, halo_windows, rate, view):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200) + world * 11 + halo_windows + (400 if view else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, halo_windows, rate, q, view)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(world + 1)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    verdict = [g for g in got if g[0] is True or g[0] is False]
    assert verdict and verdict[0][0], got
    print(got)
