"""ctypes wrapper around oracle/libnfc_oracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (see oracle/nfc_oracle.h).  It is the checker, never the
thing shipped: nothing under usrp_nfc_b200/ imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnfc_oracle.so")

EVENT_DTYPE = np.dtype([("pos", "<i8"), ("d", "<i4"), ("v", "i1"), ("type", "i1"), ("pad", "<i2")])
SYMBOL_DTYPE = np.dtype([("pos", "<i8"), ("type", "i1"), ("val", "i1"), ("pad", "<i2"), ("pad2", "<i4")])
FRAME_DTYPE = np.dtype([("pos", "<i8"), ("bit_off", "<i8"), ("nbits", "<i4"), ("type", "<i4")])
assert EVENT_DTYPE.itemsize == 16 and SYMBOL_DTYPE.itemsize == 16 and FRAME_DTYPE.itemsize == 24


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "nfc_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _ChainResult(C.Structure):
    _fields_ = [("n_events", C.c_int64), ("n_symbols", C.c_int64), ("n_frames", C.c_int64),
                ("n_bits", C.c_int64), ("digest", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.nfc_ts_new.restype = C.c_void_p
        L.nfc_ts_new.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
        L.nfc_ts_free.argtypes = [C.c_void_p]
        L.nfc_ts_work.restype = C.c_int64
        L.nfc_ts_work.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int)]
        L.nfc_ts_events.restype = C.c_void_p
        L.nfc_ts_events.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.nfc_ts_get_scalars.argtypes = [C.c_void_p, C.POINTER(C.c_double)] + [C.POINTER(C.c_int)] * 6
        L.nfc_ts_ring.restype = C.c_void_p
        L.nfc_ts_ring.argtypes = [C.c_void_p]
        L.nfc_ts_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]
        L.nfc_dec_get_state.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int]
        L.nfc_dec_set_state.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nfc_dec_new.restype = C.c_void_p
        L.nfc_dec_new.argtypes = [C.c_int, C.c_int]
        L.nfc_dec_free.argtypes = [C.c_void_p]
        L.nfc_dec_clear_outputs.argtypes = [C.c_void_p]
        L.nfc_dec_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double]
        for name in ("nfc_dec_symbols", "nfc_dec_frames", "nfc_dec_bits"):
            f = getattr(L, name)
            f.restype = C.c_void_p
            f.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.nfc_fix_ending.restype = C.c_int32
        L.nfc_fix_ending.argtypes = [C.c_void_p, C.c_int32, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.nfc_check_parity.restype = C.c_int32
        L.nfc_check_parity.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.nfc_print_enc.restype = C.c_int32
        L.nfc_print_enc.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.nfc_crc_a.restype = None
        L.nfc_crc_a.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.nfc_check_crc.restype = C.c_int
        L.nfc_check_crc.argtypes = [C.c_void_p, C.c_int32]
        L.nfc_miller_encode.restype = C.c_int32
        L.nfc_miller_encode.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.nfc_manchester_encode.restype = C.c_int32
        L.nfc_manchester_encode.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.nfc_chain_run.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(_ChainResult)]
        _lib = L
    return _lib


def _copy(ptr, count, dtype):
    if count == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (count * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


class TransitionSink:
    """transition_sink.py:10-125 (same constructor parameters and defaults)."""

    def __init__(self, samp_rate, lo_val=0.1, hi_val=1.1, av_window=2000, max_len=50):
        self._L = lib()
        self.factor = 1e6 / samp_rate
        self.av_window, self.max_len = av_window, max_len
        self._h = self._L.nfc_ts_new(samp_rate, lo_val, hi_val, av_window, max_len)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.nfc_ts_free(self._h)
            self._h = None

    def work(self, samples):
        """One work() call.  Returns (consumed, events or None); None = callback not invoked."""
        x = np.ascontiguousarray(samples, dtype=np.float32)
        cb = C.c_int(0)
        used = self._L.nfc_ts_work(self._h, x.ctypes.data, x.size, C.byref(cb))
        if not cb.value:
            return used, None
        n = C.c_int64(0)
        p = self._L.nfc_ts_events(self._h, C.byref(n))
        return used, _copy(p, n.value, EVENT_DTYPE)

    def state(self):
        ss = C.c_double()
        vals = [C.c_int() for _ in range(6)]
        self._L.nfc_ts_get_scalars(self._h, C.byref(ss), *[C.byref(v) for v in vals])
        names = ("cur_state", "dur", "last_bit", "index", "filled", "stable")
        out = {"ss": ss.value}
        out.update({k: v.value for k, v in zip(names, vals)})
        out["ring"] = _copy(self._L.nfc_ts_ring(self._h), self.av_window, np.dtype("<f8"))
        return out


    def set_state(self, ring, ss, cur_state, dur, last_bit, index, pos):
        r = np.ascontiguousarray(ring, dtype=np.float64)
        self._L.nfc_ts_set_state(self._h, r.ctypes.data, ss, cur_state, dur, last_bit, index, pos)


class Decoders:
    """background.py + manchester.py + miller.py + packets.py on the event stream."""

    def get_state(self):
        mi, ma = C.c_int(), C.c_int()
        started = np.zeros(2, np.int32)
        npend = np.zeros(2, np.int32)
        bits = np.zeros(1 << 16, np.uint8)
        self._L.nfc_dec_get_state(self._h, C.byref(mi), C.byref(ma), started.ctypes.data, npend.ctypes.data,
                                  bits.ctypes.data, bits.size)
        return dict(miller=mi.value, manch=ma.value, started=started.tolist(), pending=npend.tolist(),
                    pending_bits=bits[: int(npend.sum())].copy())

    def set_state(self, st):
        started = np.asarray(st["started"], np.int32)
        npend = np.asarray(st["pending"], np.int32)
        bits = np.ascontiguousarray(st["pending_bits"], np.uint8)
        self._L.nfc_dec_set_state(self._h, st["miller"], st["manch"], started.ctypes.data, npend.ctypes.data,
                                  bits.ctypes.data if bits.size else None)

    def __init__(self, reader=True, tag=True):
        self._L = lib()
        self._h = self._L.nfc_dec_new(int(bool(reader)), int(bool(tag)))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.nfc_dec_free(self._h)
            self._h = None

    def feed(self, events, factor):
        ev = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        self._L.nfc_dec_feed(self._h, ev.ctypes.data, ev.size, factor)

    def symbols(self):
        n = C.c_int64(0)
        return _copy(self._L.nfc_dec_symbols(self._h, C.byref(n)), n.value, SYMBOL_DTYPE)

    def frames(self):
        """Returns (frame records, list of per-frame uint8 bit arrays)."""
        n = C.c_int64(0)
        fr = _copy(self._L.nfc_dec_frames(self._h, C.byref(n)), n.value, FRAME_DTYPE)
        nb = C.c_int64(0)
        bits = _copy(self._L.nfc_dec_bits(self._h, C.byref(nb)), nb.value, np.dtype("u1"))
        return fr, [bits[f["bit_off"]: f["bit_off"] + f["nbits"]] for f in fr]

    def clear(self):
        self._L.nfc_dec_clear_outputs(self._h)


def decode_capture(samples, samp_rate, lo_val=0.1, hi_val=1.1, av_window=2000, max_len=50,
                   reader=True, tag=True, chunk=8192):
    """Run the whole path the way the flowgraph does: work() in `chunk`-item calls."""
    ts = TransitionSink(samp_rate, lo_val, hi_val, av_window, max_len)
    dec = Decoders(reader, tag)
    x = np.ascontiguousarray(samples, dtype=np.float32)
    off, evs = 0, []
    while off < x.size:
        used, ev = ts.work(x[off: off + chunk])
        if ev is not None:
            evs.append(ev)
            dec.feed(ev, ts.factor)
        off += used
    events = np.concatenate(evs) if evs else np.zeros(0, EVENT_DTYPE)
    fr, bits = dec.frames()
    return {"events": events, "symbols": dec.symbols(), "frames": fr, "frame_bits": bits, "sink": ts}


def chain_run(samples, samp_rate, lo_val, hi_val, av_window, max_len, reader=True, tag=True, chunk=8192):
    """Whole chain inside C (no Python per chunk) -- the timed CPU baseline."""
    x = np.ascontiguousarray(samples, dtype=np.float32)
    res = _ChainResult()
    lib().nfc_chain_run(x.ctypes.data, x.size, samp_rate, lo_val, hi_val, av_window, max_len,
                        int(bool(reader)), int(bool(tag)), chunk, C.byref(res))
    return {k: getattr(res, k) for k, _ in _ChainResult._fields_}


def fix_ending(bits, packet_type):
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.zeros(b.size + 1, dtype=np.uint8)
    flag = C.c_int(0)
    n = lib().nfc_fix_ending(b.ctypes.data, b.size, packet_type, out.ctypes.data, C.byref(flag))
    return out[:n].copy(), flag.value


def check_parity(bits):
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.zeros(b.size // 8 + 2, dtype=np.uint8)
    n = lib().nfc_check_parity(b.ctypes.data, b.size, out.ctypes.data)
    return None if n < 0 else out[:n].copy()


def print_enc(bits):
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.zeros(b.size // 8 + 2, dtype=np.uint8)
    fl = np.zeros(b.size // 8 + 2, dtype=np.uint8)
    n = lib().nfc_print_enc(b.ctypes.data, b.size, out.ctypes.data, fl.ctypes.data)
    return out[:n].copy(), fl[:n].copy()


def crc_a(data):
    """utilities.CRC.calculate_crc(data) -> [lo, hi]."""
    d = np.ascontiguousarray(data, dtype=np.uint8)
    out = np.zeros(2, dtype=np.uint8)
    lib().nfc_crc_a(d.ctypes.data, d.size, out.ctypes.data)
    return [int(out[0]), int(out[1])]


def check_crc(data):
    d = np.ascontiguousarray(data, dtype=np.uint8)
    return bool(lib().nfc_check_crc(d.ctypes.data, d.size))


def _encode(fn, bits):
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    lv = np.zeros(3 * b.size + 16, dtype=np.int8)
    du = np.zeros(3 * b.size + 16, dtype=np.float64)
    n = fn(b.ctypes.data, b.size, lv.ctypes.data, du.ctypes.data)
    return lv[:n].copy(), du[:n].copy()


def miller_encode(bits):
    return _encode(lib().nfc_miller_encode, bits)


def manchester_encode(bits):
    return _encode(lib().nfc_manchester_encode, bits)
