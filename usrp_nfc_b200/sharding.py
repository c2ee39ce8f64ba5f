"""Time-segment sharding of one long capture across GPUs, one process per GPU.

Rank 0 decodes its shard from the true stream start.  Every other rank starts cold `halo` samples before
its shard, exactly like the reference's warm-up (transition_sink.py:109-125): the first av_window samples
fill the ring unconditionally, the halo is decoded and thrown away.  That is speculation, so each seam is
verified: the state rank r-1 really ended in must equal the state rank r assumed at its first sample
(ring and ss bit for bit, cur_state / last_bit / dur, decoder and framer state, pending frame bits).  A rank
whose assumption was wrong decodes its shard again from its predecessor's true state.  The only traffic is
one all_gather of seam states (about 4*av_window bytes each) and a gather of frame records; there is no
data-path collective.  Frames belong to the shard in which they close; the merged log is ordered by rank,
which is the order of the closing positions.

The engine is duck-typed (push_all, state, set_state, reset, drain_frames): the product passes
usrp_nfc_b200._cabi.Stream; the gloo tests on CPU pass an oracle-backed stand-in.
"""
import numpy as np

INT64_MIN = -(2 ** 63)
PEND_CAP = 1024  # bits of an unfinished frame carried across a seam (a frame holds at most a few hundred)


def plan(total, world, av_window, halo_windows=16):
    """Shard boundaries and halo: multiples of av_window (ring slots line up) and of 4 (vector loads)."""
    import math
    q = av_window * 4 // math.gcd(av_window, 4)
    per = -(-total // world)
    per = -(-per // q) * q
    bounds = [(min(r * per, total), min((r + 1) * per, total)) for r in range(world)]
    halo = halo_windows * av_window
    halo = -(-halo // q) * q
    return bounds, halo


class SeamState(object):
    """Engine state in absolute stream coordinates as one flat byte vector (for all_gather): a header of sixteen 64-bit
    slots (positions and counters as int64, ss as the bits of its double), the ring as float32, the bits of an unfinished
    frame.  `overflow` (more than PEND_CAP open bits) travels with it, so that every rank raises after the collective."""
    NH = 16

    def __init__(self, vec, L):
        self.vec, self.L = vec, L

    @staticmethod
    def size(L):
        return SeamState.NH * 8 + L * 4 + PEND_CAP

    @staticmethod
    def zeros(L):
        return SeamState(np.zeros(SeamState.size(L), dtype=np.uint8), L)

    def _hdr(self):
        return self.vec[: self.NH * 8].view(np.int64)

    def _ring(self):
        return self.vec[self.NH * 8: self.NH * 8 + self.L * 4].view(np.float32)

    def _pend(self):
        return self.vec[self.NH * 8 + self.L * 4:]

    @staticmethod
    def from_engine(engine, base, L):
        st, ring, pend = engine.state()
        out = SeamState.zeros(L)
        h = out._hdr()
        h[0] = st.pos + base
        h[1] = np.array([st.ss], dtype=np.float64).view(np.int64)[0]
        h[2:5] = (st.cur_state, st.last_bit, st.dur)
        h[5:7] = (st.miller_state, st.manch_state)
        h[7:9] = (st.started[0], st.started[1])
        h[9:11] = (st.pending[0], st.pending[1])
        h[11] = st.serial_mode
        h[12] = 1 if len(pend) > PEND_CAP else 0
        out._ring()[:] = np.asarray(ring, dtype=np.float32)
        n = min(len(pend), PEND_CAP)
        out._pend()[:n] = np.asarray(pend[:n], dtype=np.uint8)
        return out

    def overflow(self):
        return bool(self._hdr()[12])

    def npend(self):
        h = self._hdr()
        return int(min(h[9] + h[10], PEND_CAP))

    def equal(self, other, strict_dur=False):
        a, b = self._hdr(), other._hdr()
        if a[0] != b[0] or a[1] != b[1]:  # position; ss bit for bit
            return False
        if not np.array_equal(a[2:4], b[2:4]) or not np.array_equal(a[5:11], b[5:11]):
            return False
        idle = a[2] == 0 and a[3] == 0  # dur then only phases the dropped type -1 events
        if (strict_dur or not idle) and a[4] != b[4]:
            return False
        if not np.array_equal(self._ring().view(np.uint32), other._ring().view(np.uint32)):
            return False
        n = self.npend()
        return np.array_equal(self._pend()[:n], other._pend()[:n])

    def apply(self, engine, state_cls):
        """Load into a freshly reset engine, in absolute coordinates (base 0)."""
        h, L = self._hdr(), self.L
        st = state_cls()
        st.pos = int(h[0])
        st.ss = float(np.array([h[1]], dtype=np.int64).view(np.float64)[0])
        st.cur_state, st.last_bit, st.dur = int(h[2]), int(h[3]), int(h[4])
        st.index = int(h[0]) % L
        st.stable = 1
        st.miller_state, st.manch_state = int(h[5]), int(h[6])
        st.started[0], st.started[1] = int(h[7]), int(h[8])
        st.pending[0], st.pending[1] = int(h[9]), int(h[10])
        st.serial_mode = int(h[11])
        st.lastL = INT64_MIN  # derived from (cur_state, last_bit, dur) by set_state
        st.lrun_start = INT64_MIN
        engine.set_state(st, self._ring().copy(), self._pend()[: self.npend()].copy())


def _all_gather(vec, dist, group, device):
    import torch
    world = dist.get_world_size(group)
    t = torch.from_numpy(vec).to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [o.cpu().numpy() for o in out]


def _broadcast(vec, src, dist, group, device):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(vec)).to(device)
    dist.broadcast(t, src=src, group=group)
    return t.cpu().numpy()


def _drain(engine, flat):
    if flat == "view" and hasattr(engine, "view_frames"):
        rec, b0, b1 = engine.view_frames()  # zero-copy: the caller releases them (engine.release_frames())
        return rec, (b0, b1), True
    if flat and hasattr(engine, "drain_frames_flat"):
        try:
            rec, bits = engine.drain_frames_flat(reuse=True)  # views of the engine's buffers: no allocation per shard
        except TypeError:
            rec, bits = engine.drain_frames_flat()
        return rec, bits, True
    rec, bits = engine.drain_frames()
    return rec, bits, False


def _discard(engine):
    """Frames closed inside the halo belong to the previous shard."""
    if hasattr(engine, "view_frames") and hasattr(engine, "release_frames"):
        engine.view_frames()
        engine.release_frames()
    else:
        engine.drain_frames()


def _push_range(engine, fetch, a, b, piece):
    """Samples [a, b) into the engine, at most `piece` at a time (a capture larger than the memory of the device is rendered
    or loaded piece by piece: fetch may reuse one buffer, the engine has consumed a piece when push_all returns)."""
    if not piece or piece >= b - a:
        engine.push_all(fetch(a, b))
        return
    while a < b:
        m = min(piece, b - a)
        engine.push_all(fetch(a, a + m))
        a += m


def decode_time_sharded(engine, fetch, total, av_window, state_cls, dist=None, group=None, device="cpu",
                        halo_windows=16, strict_dur=False, flat=False, piece=None):
    """Decode this rank's time shard of a `total`-sample capture.

    fetch(a, b) returns samples [a, b) (numpy array, or a CUDA tensor for a device-resident capture).
    Returns dict(frames=[(abs_pos, type, bits)], repaired=bool, seam_ok=[...], bounds=(begin, end)).
    With flat=True the frames stay in the engine's bulk form: dict(records, bits, pos_offset, n_frames); with
    flat="view" they are not even copied: bits = (bits_tag, bits_reader) as Stream.view_frames returns them, valid until
    the caller's engine.release_frames().
    piece: samples per fetch (None: the whole shard at once).
    """
    rank = dist.get_rank(group) if dist is not None else 0
    world = dist.get_world_size(group) if dist is not None else 1
    import time
    L = av_window
    bounds, halo = plan(total, world, L, halo_windows)
    begin, end = bounds[rank]
    base = 0
    tp = [time.perf_counter()]
    engine.reset()
    if rank > 0 and begin - halo - L > 0:
        base = begin - halo - L
        if begin > base:
            engine.push_all(fetch(base, begin))
        _discard(engine)
    elif rank > 0:
        engine.push_all(fetch(0, begin))  # shard too close to the stream start: decode from the true start
        _discard(engine)
    assumed = SeamState.from_engine(engine, base, L) if rank > 0 else None
    tp.append(time.perf_counter())  # halo decoded, state at the seam taken
    if end > begin:
        _push_range(engine, fetch, begin, end, piece)
    rec, bits, is_flat = _drain(engine, flat)
    tp.append(time.perf_counter())  # shard decoded, frames on the host
    pos_offset = base
    frames = None if is_flat else [(int(r["pos"]) + base, int(r["type"]), b) for r, b in zip(rec, bits)]
    final = SeamState.from_engine(engine, base, L) if begin < end or rank == 0 else assumed
    repaired, seam_ok = False, [True] * world
    if world > 1:
        zero = SeamState.zeros(L).vec
        both = _all_gather(np.concatenate([final.vec, assumed.vec if assumed is not None else zero]), dist, group, device)
        nv = SeamState.size(L)
        finals = [SeamState(np.ascontiguousarray(v[:nv]), L) for v in both]
        assumes = [SeamState(np.ascontiguousarray(v[nv:]), L) for v in both]
        if any(f.overflow() for f in finals) or any(a.overflow() for a in assumes):  # the same verdict on every rank
            raise ValueError("unfinished frame longer than PEND_CAP bits at a seam")
        for k in range(1, world):
            ok = finals[k - 1].equal(assumes[k], strict_dur)
            seam_ok[k] = ok
            if ok:
                continue
            # rank k redoes its shard from the true state; everybody learns its new final state
            if rank == k:
                engine.reset()
                finals[k - 1].apply(engine, state_cls)
                if end > begin:
                    _push_range(engine, fetch, begin, end, piece)
                rec, bits, is_flat = _drain(engine, flat)
                pos_offset = 0
                frames = None if is_flat else [(int(r["pos"]), int(r["type"]), b) for r, b in zip(rec, bits)]
                final = SeamState.from_engine(engine, 0, L)
                repaired = True
                vec = final.vec
            else:
                vec = SeamState.zeros(L).vec
            finals[k] = SeamState(_broadcast(vec, k, dist, group, device), L)
    tp.append(time.perf_counter())  # seams verified (and a wrong shard done again)
    out = dict(frames=frames, repaired=repaired, seam_ok=seam_ok, bounds=(begin, end), halo=halo, n_frames=len(rec),
               phases_ms=dict(halo=(tp[1] - tp[0]) * 1e3, shard=(tp[2] - tp[1]) * 1e3, seams=(tp[3] - tp[2]) * 1e3))
    if frames is None:
        out.update(records=rec, bits=bits, pos_offset=pos_offset)
    return out


# A frame offset in eight bytes (the layout of nfc_stream_view_frame_index, include/usrp_nfc_b200.h): position relative to the
# stream the shard was decoded as << 24 | length in bits << 8 | type.  The absolute closing position is pos_offset + position.
def pack_records(rec):
    """Frame records (fields pos, nbits, type) -> uint64 array, one packed offset per frame."""
    pos = np.asarray(rec["pos"], dtype=np.int64)
    nb = np.asarray(rec["nbits"], dtype=np.int64)
    if len(pos) and (pos.min() < 0 or pos.max() >= 1 << 40 or nb.min() < 0 or nb.max() >= 1 << 16):
        raise ValueError("a frame does not fit the packing (position >= 2^40 or more than 65535 bits)")
    u = np.uint64
    return (pos.astype(u) << u(24)) | (nb.astype(u) << u(8)) | np.asarray(rec["type"], dtype=np.int64).astype(u)


FRAME_INDEX_DTYPE = np.dtype([("pos", "<i8"), ("nbits", "<i4"), ("type", "i1"), ("shard", "i1"), ("pad", "<i2")])


def unpack_records(parts, pos_offsets):
    """Packed offsets of the shards (in shard order) -> one array with absolute closing position, length, type, shard."""
    out = np.zeros(sum(len(p) for p in parts), dtype=FRAME_INDEX_DTYPE)
    o = 0
    for r, (p, off) in enumerate(zip(parts, pos_offsets)):
        n = len(p)
        if n:
            v = np.asarray(p).view(np.uint64)
            out["pos"][o: o + n] = (v >> np.uint64(24)).astype(np.int64) + int(off)
            out["nbits"][o: o + n] = ((v >> np.uint64(8)) & np.uint64(0xffff)).astype(np.int32)
            out["type"][o: o + n] = (v & np.uint64(0xff)).astype(np.int8)
            out["shard"][o: o + n] = r
        o += n
    return out


class FrameIndex(object):
    """The frame offsets of all shards on rank 0, in stream order (shard order = closing order, packets.py:94-98): the packed
    records as they arrived (one uint64 array per shard, views of page-locked memory when they came over NCCL) and the
    position offset of every shard.  len(), positions() and unpack() decode them; nothing is decoded before it is asked for."""

    def __init__(self, parts, pos_offsets):
        self.parts, self.pos_offsets = parts, [int(o) for o in pos_offsets]

    def __len__(self):
        return sum(len(p) for p in self.parts)

    def positions(self):
        return np.concatenate([(np.asarray(p).view(np.uint64) >> np.uint64(24)).astype(np.int64) + o for p, o in zip(self.parts, self.pos_offsets)]) \
            if self.parts else np.zeros(0, dtype=np.int64)

    def unpack(self):
        return unpack_records(self.parts, self.pos_offsets)


class SharedFrameIndex(object):
    """The packed frame indexes of the ranks of ONE node in shared memory: every rank's stream writes its index straight into
    a segment of its own (nfc_stream_set_frame_index_buffer) that rank 0 maps, so that gathering the frame offsets moves no
    data -- only the counts travel (one small all_gather).  Collective: all ranks construct it with the same capacity
    (records per rank) and close() it together; it fails over to the NCCL / gloo gather when shared memory or the stream's
    call is not available (ok is False then)."""

    def __init__(self, engine, capacity, dist, group=None):
        from multiprocessing import shared_memory
        import ctypes
        self.engine, self.dist, self.group = engine, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.capacity = int(capacity)
        self.mine, self.others, self.ok = None, [], False
        name = None
        try:
            if hasattr(engine, "set_frame_index_buffer"):
                self.mine = shared_memory.SharedMemory(create=True, size=self.capacity * 8)
                self._addr = ctypes.addressof(ctypes.c_char.from_buffer(self.mine.buf))
                engine.release_frames()
                engine.set_frame_index_buffer(self._addr, self.capacity)
                name = self.mine.name
        except Exception:
            name = None
        names = [None] * self.world
        dist.all_gather_object(names, name, group=group)
        self.ok = all(n is not None for n in names)
        if self.ok and self.rank == 0:
            try:
                self.others = [self.mine if r == 0 else shared_memory.SharedMemory(name=names[r]) for r in range(self.world)]
            except Exception:
                self.ok = False
        flags = [None] * self.world
        dist.all_gather_object(flags, self.ok, group=group)
        self.ok = all(flags)
        if not self.ok:
            self._detach()

    def parts(self, counts):
        """rank 0: the ranks' packed indexes as uint64 views of the shared segments (valid until a rank decodes again)."""
        return [np.frombuffer(self.others[r].buf, dtype=np.uint64, count=int(counts[r])) for r in range(self.world)]

    def _detach(self):
        if self.mine is not None:
            try:
                self.engine.release_frames()
                self.engine.set_frame_index_buffer(None, 0)
            except Exception:
                pass
        for m in self.others:
            if m is not self.mine:
                try:
                    m.close()
                except Exception:
                    pass
        self.others = []
        if self.mine is not None:
            try:
                del self._addr
                self.mine.close()
                self.mine.unlink()
            except Exception:
                pass
            self.mine = None

    def close(self):
        if self.dist is not None and self.ok:
            self.dist.barrier(group=self.group)  # rank 0 is done reading
        self.ok = False
        self._detach()


def gather_frame_records(engine_or_rec, pos_offset, dist=None, group=None, device="cpu", state=None, shared=None):
    """The frame offsets of all shards on rank 0 in stream order (the order the reference hands frames to fsm.process_bits,
    packets.py:94-98).  One all_gather of (count, position offset); then either nothing more -- `shared`: a SharedFrameIndex,
    the ranks of one node keep their packed indexes in shared memory that rank 0 maps -- or one gather of the packed records
    (8 bytes per frame; device tensors over NCCL, fed from and read back into page-locked memory, or host tensors over gloo).
    engine_or_rec: a Stream (its packed index is used as it lies: nfc_stream_view_frame_index) or an array of frame records.
    Returns a FrameIndex on rank 0, None elsewhere.  `state`: a dict that keeps the staging tensors between calls."""
    import torch
    if hasattr(engine_or_rec, "view_frame_index"):
        packed = engine_or_rec.view_frame_index()
    else:
        packed = pack_records(engine_or_rec)
    if dist is None or dist.get_world_size(group) == 1:
        return FrameIndex([packed], [pos_offset])
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    state = state if state is not None else {}
    on_gpu = str(device) != "cpu"
    cnt = torch.tensor([len(packed), int(pos_offset)], dtype=torch.int64, device=device)
    cnts = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    meta = torch.stack(cnts).cpu().numpy()
    if shared is not None and shared.ok:
        # every rank's index is complete in its segment (view_frame_index settled the stream before the all_gather)
        return FrameIndex(shared.parts(meta[:, 0]), meta[:, 1]) if rank == 0 else None
    cap = max(1, int(meta[:, 0].max()))
    if state.get("cap", 0) < cap:
        state["cap"] = int(cap * 1.25) + 16
        state["send"] = torch.zeros(state["cap"], dtype=torch.int64, device=device)
        state["recv"] = [torch.zeros(state["cap"], dtype=torch.int64, device=device) for _ in range(world)] if rank == 0 else None
        state["host"] = torch.zeros((world, state["cap"]), dtype=torch.int64).pin_memory() if (rank == 0 and on_gpu) else None
    send = state["send"]
    if len(packed):
        send[: len(packed)].copy_(torch.from_numpy(packed.view(np.int64)), non_blocking=True)
    dist.gather(send, state["recv"] if rank == 0 else None, dst=0, group=group)
    if on_gpu:
        torch.cuda.current_stream().synchronize()  # the engine may reuse its index once this returns
    if rank != 0:
        return None
    if state["host"] is not None:
        for r in range(world):
            state["host"][r, : int(meta[r, 0])].copy_(state["recv"][r][: int(meta[r, 0])], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        parts = [state["host"][r, : int(meta[r, 0])].numpy().view(np.uint64) for r in range(world)]
    else:
        parts = [state["recv"][r][: int(meta[r, 0])].numpy().view(np.uint64).copy() for r in range(world)]
    return FrameIndex(parts, meta[:, 1])


def gather_frames(frames, dist=None, group=None):
    """Frame records of all ranks on rank 0, in stream order (rank order = closing-position order)."""
    if dist is None or dist.get_world_size(group) == 1:
        return frames
    out = [None] * dist.get_world_size(group) if dist.get_rank(group) == 0 else None
    payload = [(p, t, np.asarray(b, dtype=np.uint8).tobytes()) for p, t, b in frames]
    dist.gather_object(payload, out, dst=0, group=group)
    if out is None:
        return None
    merged = []
    for part in out:
        merged.extend((p, t, np.frombuffer(b, dtype=np.uint8)) for p, t, b in part)
    return merged
