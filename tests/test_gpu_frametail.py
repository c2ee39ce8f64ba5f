"""The per-frame tail on the device (csrc/frametail.cu) against the reference's fsm / utilities known answers and the oracle.

Reference: fsm._fix_ending (fsm.py:51-66), _check_parity (:28-49), _print_enc (:114-131), utilities.CRC (utilities.py:26-46).
"""
import numpy as np
import pytest

from oracle import oracle
from tests import helpers as H
from usrp_nfc_b200 import _cabi, synth

pytestmark = pytest.mark.gpu


def _pack(frames_bits, types, split=True):
    """[(bits)], [type] -> FRAME_DTYPE records + the two bit buffers view_frames would hand out."""
    fr = np.zeros(len(frames_bits), dtype=_cabi.FRAME_DTYPE)
    bufs = [[], []]
    for i, (bits, t) in enumerate(zip(frames_bits, types)):
        which = t if split else 0
        fr[i] = (i, len(bufs[which]), len(bits), t)
        bufs[which].extend(int(b) for b in bits)
    b0 = np.array(bufs[0], dtype=np.uint8)
    b1 = np.array(bufs[1], dtype=np.uint8)
    return fr, b0, (b1 if split else None)


def _check_against_oracle(frames_bits, types, tails, by, fl):
    for i, (bits, t) in enumerate(zip(frames_bits, types)):
        fixed, flag = oracle.fix_ending(bits, t)
        tl = tails[i]
        assert int(tl["nbits"]) == fixed.size and int(tl["fix_flag"]) == flag, i
        eb, ef = oracle.print_enc(fixed)
        o, n = int(tl["byte_off"]), int(tl["nbytes"])
        assert n == eb.size and by[o:o + n].tolist() == eb.tolist() and fl[o:o + n].tolist() == ef.tolist(), i
        par = oracle.check_parity(fixed)
        ok = par is not None and par.size > 0
        assert bool(tl["parity_ok"]) == ok, i
        if ok:
            assert par.tolist() == by[o:o + n].tolist()
        assert bool(tl["crc_ok"]) == (oracle.check_crc(eb) if n >= 2 else False), i


def test_golden_fsm_tail():
    recs = H.load_json("fsm_tail.json")
    fr, b0, b1 = _pack([r["bits"] for r in recs], [r["type"] for r in recs])
    tails, by, fl = _cabi.frames_tail(fr, b0, b1)
    for r, tl in zip(recs, tails):
        assert int(tl["nbits"]) == len(r["fixed"])
        assert {0: "", 1: "EXTRA ERROR", 2: "MANY MORE ERROR"}[int(tl["fix_flag"])] == r["msg"]
        o, n = int(tl["byte_off"]), int(tl["nbytes"])
        assert [[int(b), int(f)] for b, f in zip(by[o:o + n], fl[o:o + n])] == r["enc"]
        want_ok = r["parity"] is not None and len(r["parity"]) > 0
        assert bool(tl["parity_ok"]) == want_ok
        if want_ok:
            assert by[o:o + n].tolist() == r["parity"]


def test_golden_crc_a():
    cases = H.load_json("crc_a.json")
    frames, types, want = [], [], []
    for k, rec in enumerate(cases):
        for payload, ok in ((rec["data"] + rec["crc"], rec["check_good"]), (rec["bad"], rec["check_bad"])):
            frames.append(synth.bytes_to_bits(payload))
            types.append(k & 1)
            want.append(ok)
    fr, b0, b1 = _pack(frames, types)
    tails, by, fl = _cabi.frames_tail(fr, b0, b1)
    assert [bool(c) for c in tails["crc_ok"]] == want
    assert all(bool(p) for p in tails["parity_ok"]) and not fl.any()


def test_random_frames_match_oracle_single_buffer():
    rng = np.random.default_rng(5)
    frames, types = [], []
    for i in range(6000):
        kind = i % 4
        if kind == 0:  # well-formed, sometimes a bit short or long at the end (what the framer really produces)
            payload = rng.integers(0, 256, int(rng.integers(1, 20))).tolist()
            if rng.random() < 0.5 and len(payload) >= 1:
                payload = payload + oracle.crc_a(payload)
            bits = synth.bytes_to_bits(payload)
            cut = int(rng.integers(0, 3))
            bits = bits[:len(bits) - cut] + rng.integers(0, 2, int(rng.integers(0, 2))).tolist()
        elif kind == 1:
            bits = rng.integers(0, 2, int(rng.integers(0, 200))).tolist()
        elif kind == 2:
            bits = rng.integers(0, 2, int(rng.integers(0, 12))).tolist()
        else:
            bits = [int(rng.integers(0, 2))] * int(rng.integers(0, 40))
        frames.append(bits)
        types.append(int(rng.integers(0, 2)))
    fr, b0, _ = _pack(frames, types, split=False)
    tails, by, fl = _cabi.frames_tail(fr, b0)  # one buffer for both types, as drain_frames_flat returns it
    _check_against_oracle(frames, types, tails, by, fl)
    assert int(tails["byte_off"][-1]) + int(tails["nbytes"][-1]) == by.size


def test_decoded_stream_frames():
    """Frames the device decoded from rendered traffic: plain traffic passes parity and CRC end to end."""
    sess = synth.load_sessions()
    pcm = synth.capture(sess["ultralight"], 2e6, 31, sessions=2)
    x = synth.envelope(synth.pcm_to_float(pcm))
    s = _cabi.Stream(2e6, hi_val=1.09)
    s.push_all(x)
    s.push_all(np.full(4000, float(np.mean(x[:1000])), np.float32))
    fr, b0, b1 = s.view_frames()
    tails, by, fl = _cabi.frames_tail(fr, b0, b1)
    frames = [(b0 if int(r["type"]) == 0 else b1)[int(r["bit_off"]):int(r["bit_off"]) + int(r["nbits"])].tolist() for r in fr]
    _check_against_oracle(frames, [int(r["type"]) for r in fr], tails, by, fl)
    s.release_frames()
    s.close()
    assert len(fr) > 10 and int(tails["parity_ok"].sum()) >= len(fr) - 2


def test_empty_batch_and_bad_records():
    tails, by, fl = _cabi.frames_tail(np.zeros(0, dtype=_cabi.FRAME_DTYPE), np.zeros(0, np.uint8))
    assert tails.size == 0 and by.size == 0
    fr = np.zeros(1, dtype=_cabi.FRAME_DTYPE)
    fr[0] = (0, 5, 20, 0)
    with pytest.raises(_cabi.NfcError):
        _cabi.frames_tail(fr, np.zeros(10, np.uint8))
