"""Executable specification of the PARALLEL formulation the CUDA kernels implement -- TEST INFRASTRUCTURE.

The reference's slicer is a per-sample loop with state (ss, ring, cur_state, last_bit, dur).
The device path never runs that loop; it derives the same outputs from position-indexed
quantities.  This module states those derivations in plain Python/numpy so that they can be
checked against the oracle on the CPU (tests/test_algomodel.py) before any GPU is involved:

  classes  -> vals          : the hysteresis "ratio > hi is ignored while cur_state == 2" as a
                              function of the distance to the last LOW sample        (slicer.cu)
  vals     -> transitions   : positions where val changes                             (slicer.cu)
  transitions -> events     : timeouts, durations, v and type per run, O(1) lookback  (runs.cu)

Nothing here is used by the product.
"""
import numpy as np

LOW, MID, HIGH = -1, 0, 1


def classes_sequential(x, L, lo, hi, mx):
    """Per-sample ratio class with the TRUE running sum (reference recurrence, transition_sink.py:55-82).
    Returns classes (the ratio tests only; hysteresis not applied) for the stable region and the
    stream index of the first stable sample."""
    x = np.asarray(x, dtype=np.float32).astype(np.float64)
    n = x.size
    if n < L:
        return np.zeros(0, np.int8), n
    ar = list(x[:L])
    ss = 0
    for v in ar:
        ss = ss + v
    index, st = 0, 0
    dur, last_bit = L % mx, 0
    cls = np.zeros(n - L, np.int8)
    for i in range(L, n):
        bit = x[i]
        prev = ar[index]
        if ss == 0:
            ratio = 1 if bit == 0 else hi + 0.1
        else:
            ratio = bit * L / ss
        c = LOW if lo > ratio else (HIGH if ratio > hi else MID)
        cls[i - L] = c
        if c == LOW:
            val, cur, st = -1, prev, 2
        elif st != 2 and c == HIGH:
            val, cur, st = 1, prev, 1
        else:
            val, cur = 0, bit
        ar[index] = cur
        index = (index + 1) % L
        ss += cur - prev
        if val == last_bit:
            dur += 1
        else:
            dur, last_bit = 1, val
        if dur > mx:
            dur, st = 1, 0
    return cls, L


def vals_from_classes(cls, mx, lastL=None, lrun_start=None):
    """Hysteresis without a state machine.

    A HIGH-class sample at i is forced to val 0 iff cur_state == 2 when it arrives, i.e. iff
      * there is a LOW sample b < i with i - b <= mx + 1 (the 0-run after b has not timed out), and
      * b itself was not a timeout sample: with a = first sample of the LOW run containing b,
        NOT ((b - a) >= mx and (b - a) % mx == 0).
    lastL / lrun_start carry (b, a) in from before the block (None = no LOW sample yet).
    """
    n = cls.size
    val = cls.astype(np.int8).copy()
    b, a = lastL, lrun_start
    for i in range(n):
        c = cls[i]
        if c == LOW:
            if b is None or b != i - 1:
                a = i
            b = i
        elif c == HIGH:
            if b is not None and i - b <= mx + 1:
                j = b - a
                tmo = j >= mx and j % mx == 0
                if not tmo:
                    val[i] = 0
    return val, b, a


def transitions(val, last_bit):
    """Positions (relative to the block) where val differs from its predecessor, and the new val."""
    prev = np.concatenate(([last_bit], val[:-1])).astype(np.int8)
    pos = np.nonzero(val != prev)[0]
    return pos.astype(np.int64), val[pos]


def events_from_transitions(tpos, tval, w0, w1, st0, last_bit0, dur0, mx, keep_dropped=True):
    """Per-run derivation of the event stream for stream positions [w0, w1).

    tpos/tval: transitions inside the window (absolute positions, ascending).
    (st0, last_bit0, dur0): the reference's (cur_state, last_bit, dur) at w0.
    Returns (events [(pos, v, d, type)], (st, last_bit, dur) at w1).
    """
    nrun = len(tpos) + 1
    u = [int(last_bit0)] + [int(v) for v in tval]  # val of run r
    p = [w0 - dur0] + [int(q) for q in tpos]  # (virtual) start of run r
    e = [int(q) for q in tpos] + [w1]  # end (exclusive) of run r

    def tmo_at_last(r):
        ell = e[r] - p[r]
        return (ell - 1) >= mx and (ell - 1) % mx == 0

    def first_tmo_in_window(r):
        """first timeout position q = p + k*mx (k >= 1) with q >= w0 (for run 0) and q < e[r], or None"""
        k = 1
        if r == 0:
            k = max(1, -(-(w0 - p[0]) // mx))  # ceil((w0 - p0) / mx)
        q = p[r] + k * mx
        return q if q < e[r] else None

    memo = {}

    def s_end(r):
        """cur_state after the last sample of run r (before the sample that closes it)."""
        if r in memo:
            return memo[r]
        if r == 0 and e[0] == w0:
            out = st0
        elif u[r] != 0:
            out = 0 if tmo_at_last(r) else (2 if u[r] == -1 else 1)
        else:
            if first_tmo_in_window(r) is not None:
                out = 0
            else:
                out = st0 if r == 0 else s_end(r - 1)
        memo[r] = out
        return out

    ev = []
    for r in range(nrun):
        s_begin = st0 if r == 0 else s_end(r - 1)
        q = first_tmo_in_window(r)
        first = True
        while q is not None and q < e[r]:
            if u[r] == -1:
                st_q = 2
            elif u[r] == 1:
                st_q = 1
            else:
                st_q = s_begin if first else 0
            first = False
            if keep_dropped or st_q != 0:
                ev.append((q, u[r] + (1 if st_q == 2 else 0), mx, st_q - 1))
            q += mx
        if r + 1 < nrun:  # closing transition at i = e[r]
            i = e[r]
            un = u[r + 1]
            se = s_end(r)
            st_after = 2 if un == -1 else (1 if un == 1 else se)
            ell = i - p[r]
            d = mx if se == 0 else ((ell - 1) % mx) + 1
            if keep_dropped or st_after != 0:
                ev.append((i, u[r] + (1 if st_after == 2 else 0), d, st_after - 1))
    # carry out
    r = nrun - 1
    if w1 > w0:
        st1 = s_end(r)
        dur1 = ((w1 - 1 - p[r]) % mx) + 1 if mx > 0 else 1
        lb1 = u[r]
    else:
        st1, dur1, lb1 = st0, dur0, last_bit0
    return ev, (st1, lb1, dur1)


# ---- frame counts of the line-code kernels (linecode.cu: ChunkCnt, CountSink, CombineCnt, resolved_records) ----------------
# A symbol is (type, val): val 0/1 a bit, anything else closes the frame of that type if one has started
# (PacketProcessor.append_bit, packets.py:67-79); "cap" marks a capture boundary of a batch (nothing pending behind it).
def framer_items(symbols, started):
    """[(kind, type)]: the calls the device's sinks see -- 'bit' and 'close' per symbol (framer_put), given the framers'
    started flags at the beginning."""
    st = list(started)
    out = []
    for t, v in symbols:
        if t == "cap":
            out.append(("cap", 0))
            st = [False, False]
            continue
        if v in (0, 1):
            if not st[t] and v == (1 if t == 0 else 0):
                st[t] = True
            else:
                out.append(("bit", t))
        elif st[t]:
            out.append(("close", t))
            st[t] = False
    return out


def chunk_cnt(items):
    """CountSink over one chunk: nemit, has[2] (bit 0: closes / capture end, bit 1: first closing depends on what is pending
    before the chunk), tail[2]."""
    c = dict(nemit=0, has=[0, 0], tail=[0, 0], nbit=[0, 0])
    for kind, t in items:
        if kind == "bit":
            c["nbit"][t] += 1
            c["tail"][t] += 1
        elif kind == "close":
            if c["tail"][t] != 0:
                c["nemit"] += 1
            elif not (c["has"][t] & 1):
                c["nemit"] += 1
                c["has"][t] |= 2
            c["has"][t] |= 1
            c["tail"][t] = 0
        else:
            for u in (0, 1):
                c["has"][u] |= 1
                c["tail"][u] = 0
    return c


def combine_cnt(a, b):
    c = dict(nemit=a["nemit"] + b["nemit"], has=[0, 0], tail=[0, 0], nbit=[a["nbit"][0] + b["nbit"][0], a["nbit"][1] + b["nbit"][1]])
    for t in (0, 1):
        if a["has"][t] & 1:
            if (b["has"][t] & 2) and a["tail"][t] == 0:
                c["nemit"] -= 1
            c["has"][t] = a["has"][t]
        else:
            c["has"][t] = (b["has"][t] & 1) | (2 if (b["has"][t] & 2) and a["tail"][t] == 0 else 0)
        c["tail"][t] = b["tail"][t] if (b["has"][t] & 1) else a["tail"][t] + b["tail"][t]
    return c


def resolved_records(pc, pending):
    return pc["nemit"] - sum(1 for t in (0, 1) if (pc["has"][t] & 2) and pending[t] == 0)


def frames_sequential(items, pending):
    """What the write pass does: lengths of the frames that get a record, closings without a bit, bits pending at the end."""
    pend = list(pending)
    lens, empty = [], 0
    for kind, t in items:
        if kind == "bit":
            pend[t] += 1
        elif kind == "close":
            if pend[t] == 0:
                empty += 1
            else:
                lens.append((t, pend[t]))
            pend[t] = 0
        else:
            pend = [0, 0]
    return lens, empty, pend
