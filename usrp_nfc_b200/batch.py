"""Batches of independent captures (BASELINE.json configs[3]).

Every capture is its own `transition_sink` + decoders + packet processors in the reference (separate state,
its own `hi_val`), so a batch partitions trivially: captures are dealt round-robin to the ranks (one process per
GPU) and, inside a rank, to a few worker threads that each drive one `nfc_stream` at a time on its own CUDA
stream -- the kernels of different captures overlap on the device, the C ABI releases the GIL.  There is no
data-path collective; only frame counts / records are gathered.
"""
import threading

import numpy as np

from . import _cabi


def _decode_one(x, samp_rate, params, device, tuning, outputs, kind, cache, blocking=True):
    """One capture through one nfc_stream; a worker keeps one stream per window geometry and only changes its thresholds
    between captures (device buffers are not reallocated)."""
    key = tuple(sorted((k, v) for k, v in params.items() if k not in ("lo_val", "hi_val")))
    s = cache.get(key)
    if s is None:
        s = _cabi.Stream(samp_rate, device=device, outputs=outputs, input_kind=kind, **params)
        if tuning:
            s.set_tuning(**tuning)
        s.set_wait_mode(blocking)
        cache[key] = s
    else:
        s.reset()
        s.set_thresholds(params.get("lo_val", 0.1), params.get("hi_val", 1.1))
    s.push_all(x)
    fr, bits = s.drain_frames_flat()
    return fr, bits


def decode_batch(captures, samp_rate, params, device=0, workers=8, tuning=None, outputs=_cabi.OUT_FRAMES,
                 kind=_cabi.IN_ENVELOPE_F32, blocking_wait=True):
    """Decode independent captures on one GPU.

    captures: sequence of sample arrays (numpy, or CUDA tensors on `device`); params: one dict of Stream keyword
    arguments per capture (hi_val, av_window, max_len, ...) or a single dict for all.
    blocking_wait: the workers sleep while they wait for their stream (nfc_stream_set_wait_mode) instead of spinning --
    with spinning waits, the workers and each stream's marshalling thread outnumber the host cores.
    Returns a list of (frame records, flat frame bits) in capture order.
    """
    n = len(captures)
    plist = params if isinstance(params, (list, tuple)) else [params] * n
    out = [None] * n
    errors = []
    lock = threading.Lock()
    nxt = [0]

    def work():
        cache = {}
        try:
            _work(cache)
        finally:
            for st in cache.values():
                st.close()

    def _work(cache):
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= n:
                return
            try:
                out[i] = _decode_one(captures[i], samp_rate, plist[i], device, tuning, outputs, kind, cache, blocking_wait)
            except Exception as exc:  # surfaced to the caller below
                errors.append((i, exc))
                return

    ths = [threading.Thread(target=work) for _ in range(max(1, min(workers, n)))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errors:
        raise errors[0][1]
    return out


def rank_share(n_captures, rank, world):
    """Indices of the captures rank `rank` decodes (round-robin)."""
    return list(range(rank, n_captures, world))


def decode_batch_distributed(make_capture, n_captures, samp_rate, params_of, dist=None, group=None, device=0,
                             workers=8, tuning=None, kind=_cabi.IN_ENVELOPE_F32):
    """Round-robin a batch over the ranks.  make_capture(i) -> samples of capture i (called only for this rank's
    share), params_of(i) -> its Stream keyword arguments.  Returns dict(indices, results, n_frames_total)."""
    rank = dist.get_rank(group) if dist is not None else 0
    world = dist.get_world_size(group) if dist is not None else 1
    mine = rank_share(n_captures, rank, world)
    caps = [make_capture(i) for i in mine]
    res = decode_batch(caps, samp_rate, [params_of(i) for i in mine], device=device, workers=workers, tuning=tuning,
                       kind=kind)
    n_frames = sum(len(fr) for fr, _ in res)
    total = n_frames
    if dist is not None and world > 1:
        import torch
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.tensor([float(n_frames)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, group=group)
        total = int(t.item())
    return dict(indices=mine, results=res, n_frames=n_frames, n_frames_total=total)


def frames_as_lists(fr, bits):
    """(records, flat bits) -> [(pos, type, uint8 bit array)] like Stream.drain_frames."""
    return [(int(r["pos"]), int(r["type"]), np.array(bits[int(r["bit_off"]): int(r["bit_off"]) + int(r["nbits"])], dtype=np.uint8))
            for r in fr]
