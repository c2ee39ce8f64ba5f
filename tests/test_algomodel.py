"""The position-indexed (parallel) formulation of the slicer equals the oracle's per-sample loop."""
import json

import numpy as np
import pytest

from oracle import oracle
from tests import algomodel as M
from tests import helpers as H


def _model_events(x, L, lo, hi, mx, splits):
    cls, t0 = M.classes_sequential(x, L, lo, hi, mx)
    val, _, _ = M.vals_from_classes(cls, mx)
    # blockwise hysteresis with carry must agree with one pass
    b = a = None
    parts = []
    edges = [0] + sorted(set(int(s) for s in splits if 0 < s < cls.size)) + [cls.size]
    for s, e in zip(edges[:-1], edges[1:]):
        v, b2, a2 = M.vals_from_classes(cls[s:e], mx, None if b is None else b - s, None if a is None else a - s)
        b = None if b2 is None else b2 + s
        a = None if a2 is None else a2 + s
        parts.append(v)
    assert np.array_equal(np.concatenate(parts) if parts else val, val)
    tpos, tval = M.transitions(val, 0)
    tpos = tpos + t0
    # windows with carried (st, last_bit, dur)
    st, lb, dur = 0, 0, L % mx
    out = []
    wedges = [t0] + [t0 + s for s in edges[1:-1]] + [t0 + cls.size]
    for w0, w1 in zip(wedges[:-1], wedges[1:]):
        sel = (tpos >= w0) & (tpos < w1)
        ev, (st, lb, dur) = M.events_from_transitions(tpos[sel], tval[sel], w0, w1, st, lb, dur, mx)
        out.extend(ev)
    return out, (st, lb, dur)


def _oracle_events(x, L, lo, hi, mx):
    ts = oracle.TransitionSink(2e6, lo, hi, L, mx)
    off, evs = 0, []
    while off < x.size:
        used, ev = ts.work(x[off: off + 4096])
        if ev is not None:
            evs.append(ev)
        off += used
    ev = np.concatenate(evs) if evs else np.zeros(0, oracle.EVENT_DTYPE)
    return ev, ts.state()


def _compare(x, L, lo, hi, mx, splits):
    got, (st, lb, dur) = _model_events(x, L, lo, hi, mx, splits)
    want, state = _oracle_events(x, L, lo, hi, mx)
    got = np.array(got, dtype=np.int64).reshape(-1, 4)
    w = np.stack([want["pos"], want["v"], want["d"], want["type"]], 1).astype(np.int64).reshape(-1, 4)
    assert got.shape == w.shape, (got.shape, w.shape)
    bad = np.nonzero((got != w).any(axis=1))[0]
    assert bad.size == 0, "first mismatch at event %d: got %s want %s" % (bad[0], got[bad[0]], w[bad[0]])
    if state["stable"] and x.size > L:
        assert (st, lb, dur) == (state["cur_state"], state["last_bit"], state["dur"])


def test_model_on_slicer_kats():
    z = H.load_case("slicer_kat")
    meta = json.loads(bytes(z["meta"]).decode())
    rng = np.random.default_rng(5)
    for ci, m in enumerate(meta):
        x = z["x%d" % ci]
        if np.isnan(x).any() or x.size <= m["L"]:
            continue
        _compare(x, m["L"], m["lo"], m["hi"], m["mx"], rng.integers(1, max(2, x.size), 4))


@pytest.mark.parametrize("name", ["surrogate_ultralight", "rate_1356"])
def test_model_on_captures(name):
    case = H.load_case(name)
    p = H.summary()[name]
    L, mx = p.get("av_window", 2000), p.get("max_len", 50)
    x = H.case_input(case)
    _compare(x, L, 0.1, 1.09, mx, [1000, 5000, 5001, 20000])


def test_model_hysteresis_corner_cases():
    """Long pauses (timeouts inside a LOW run), spikes right after a pause, HIGH directly after LOW."""
    rng = np.random.default_rng(3)
    L, mx = 64, 7
    for trial in range(200):
        n = 600
        x = np.full(n, 0.25, np.float32)
        i = L + 5
        while i < n - 40:
            kind = rng.integers(0, 5)
            ln = int(rng.choice([1, 2, mx - 1, mx, mx + 1, 2 * mx, 2 * mx + 1, 3 * mx + 1]))
            if kind == 0:
                x[i:i + ln] = 1e-4  # pause
                i += ln
                if rng.random() < 0.7:
                    k = int(rng.integers(0, mx + 4))
                    x[i + k: i + k + int(rng.integers(1, 4))] = 0.4  # spike after the pause
            elif kind == 1:
                x[i:i + ln] = 0.4
                i += ln
            i += int(rng.integers(1, 3 * mx))
        _compare(x, L, 0.1, 1.1, mx, rng.integers(1, n, 3))


def test_frame_record_counts_leave_out_empty_frames():
    """linecode.cu ChunkCnt / CombineCnt: the scanned counts give every chunk the index of its first frame record with the
    frames that close without a bit left out (packets.py:97 does not forward them) -- whatever the chunking, whatever is
    pending when the slab begins, with capture boundaries of a batch in between.  Model of the device code, checked against
    the sequential rule on random symbol streams (the device kernels themselves: tests/test_gpu_parity.py, decoder KATs)."""
    import functools
    rng = np.random.default_rng(11)
    zero = dict(nemit=0, has=[0, 0], tail=[0, 0], nbit=[0, 0])
    for trial in range(300):
        n = int(rng.integers(0, 120))
        syms = []
        for _ in range(n):
            r = rng.random()
            if r < 0.03:
                syms.append(("cap", 0))
            else:
                syms.append((int(rng.integers(0, 2)), int(rng.choice([0, 1, 2], p=[0.35, 0.35, 0.3]))))
        started = [bool(rng.integers(0, 2)), bool(rng.integers(0, 2))]
        pending = [int(rng.integers(0, 3)) * int(rng.integers(0, 2)), int(rng.integers(0, 3)) * int(rng.integers(0, 2))]
        items = M.framer_items(syms, started)
        lens, empty, pend_end = M.frames_sequential(items, pending)
        chunk = int(rng.integers(1, 9))
        chunks = [items[i:i + chunk] for i in range(0, len(items), chunk)]
        cnts = [M.chunk_cnt(c) for c in chunks]
        prefix = zero
        written = 0
        for c, items_c in zip(cnts, chunks):
            assert M.resolved_records(prefix, pending) == written, (trial, chunk)
            pend_c = [prefix["tail"][t] if (prefix["has"][t] & 1) else pending[t] + prefix["tail"][t] for t in (0, 1)]
            lens_c, _, _ = M.frames_sequential(items_c, pend_c)
            written += len(lens_c)
            prefix = M.combine_cnt(prefix, c)
        assert written == len(lens) and M.resolved_records(prefix, pending) == len(lens)
        assert M.resolved_records(prefix, pending) <= prefix["nemit"] <= len(lens) + 2  # the scan's total is an upper bound
        # associativity: any bracketing of the chunks gives the same total
        if len(cnts) >= 3:
            k = int(rng.integers(1, len(cnts) - 1))
            left = functools.reduce(M.combine_cnt, cnts[:k], zero)
            right = functools.reduce(M.combine_cnt, cnts[k:])
            assert M.combine_cnt(left, right) == prefix
