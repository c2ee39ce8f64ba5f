"""Pin the CPU oracle (oracle/nfc_oracle.c) against fixtures produced by the reference itself
(oracle/gen_golden.py ran the reference's own modules from /root/reference/code) and against
the reference's golden logs (outputs/1k_with_enc.out, outputs/ultralight.out)."""
import json

import numpy as np
import pytest

from oracle import oracle
from tests import helpers as H


@pytest.mark.parametrize("name,rate,kw", [
    ("surrogate_classic1k", 2e6, dict(hi_val=1.09)),
    ("surrogate_ultralight", 2e6, dict(hi_val=1.09)),
    ("rate_1356", 13.56e6, dict(hi_val=1.09, av_window=13560, max_len=339)),
    ("rate_2000", 20e6, dict(hi_val=1.09, av_window=20000, max_len=500)),
])
@pytest.mark.parametrize("chunk", [8192, 1000003, 777])
def test_capture_matches_reference(name, rate, kw, chunk):
    case = H.load_case(name)
    out = oracle.decode_capture(H.case_input(case), rate, chunk=chunk, **kw)
    H.assert_events_equal(out["events"], case["ev"])
    H.assert_symbols_equal(out["symbols"], case["sym"])
    H.assert_frames_equal(out["frames"], out["frame_bits"], case)
    st = out["sink"].state()
    assert st["ss"] == float(case["state_ss"])
    assert (st["cur_state"], st["dur"], st["last_bit"], st["index"]) == tuple(
        int(case["state_" + k]) for k in ("cur_state", "dur", "last_bit", "index"))
    assert np.array_equal(st["ring"], case["state_ring"])


@pytest.mark.parametrize("key", ["classic1k", "ultralight"])
def test_frame_log_matches_reference_outputs(key):
    """Decoded frames, repaired and parity-checked, reproduce outputs/*.out: the raw encrypted bytes
    with their '!' flags where the log prints them, the plaintext bytes elsewhere."""
    case = H.load_case("surrogate_" + key)
    logged = H.load_json("logged_frames.json")[key]
    out = oracle.decode_capture(H.case_input(case), 2e6, hi_val=1.09)
    assert len(out["frames"]) == len(logged)
    for fr, bits, want in zip(out["frames"], out["frame_bits"], logged):
        assert int(fr["type"]) == want["type"]
        fixed, _ = oracle.fix_ending(bits, int(fr["type"]))
        if want["raw"] is not None:
            by, fl = oracle.print_enc(fixed)
            assert [[int(b), int(f)] for b, f in zip(by, fl)] == want["raw"]
        else:
            by = oracle.check_parity(fixed)
            assert by is not None and by.tolist() == want["bytes"]


def test_slicer_known_answers():
    z = H.load_case("slicer_kat")
    meta = json.loads(bytes(z["meta"]).decode())
    for ci, m in enumerate(meta):
        x = z["x%d" % ci]
        ts = oracle.TransitionSink(2e6, m["lo"], m["hi"], m["L"], m["mx"])
        off, k, evs, nb = 0, 0, [], []
        while off < x.size:
            n = m["chunks"][k % len(m["chunks"])]
            k += 1
            used, ev = ts.work(x[off: off + n])
            if ev is not None:
                evs.append(ev)
                nb.append(len(ev))
            off += used
        ev = np.concatenate(evs) if evs else np.zeros(0, oracle.EVENT_DTYPE)
        want = z["ev%d" % ci]
        assert nb == z["nb%d" % ci].tolist(), "case %d: per-callback batch sizes" % ci
        got = np.stack([ev["v"], ev["d"], ev["type"]], axis=1).astype(np.int32) if len(ev) else np.zeros((0, 3), np.int32)
        assert np.array_equal(got, want), "case %d: events" % ci
        st = ts.state()
        assert bool(st["stable"]) == m["stable"]
        if m["stable"]:
            same_ss = st["ss"] == m["ss"] or (np.isnan(st["ss"]) and np.isnan(m["ss"]))
            assert same_ss, "case %d: ss %r != %r" % (ci, st["ss"], m["ss"])
            assert (st["cur_state"], st["dur"], st["last_bit"], st["index"]) == (
                m["cur_state"], m["dur"], m["last_bit"], m["index"]), "case %d" % ci
        assert np.array_equal(st["ring"], z["ring%d" % ci], equal_nan=True), "case %d: ring" % ci


def test_decoder_known_answers():
    z = H.load_case("decoder_kat")
    for ci in range(40):
        evs = z["ev%d" % ci]
        ev = np.zeros(len(evs), oracle.EVENT_DTYPE)
        ev["v"], ev["d"], ev["type"] = evs[:, 0], evs[:, 1], evs[:, 2]
        ev["pos"] = np.arange(len(evs))
        dec = oracle.Decoders(True, True)
        # split into several callback batches: grouping must not matter (background.py:42-52)
        for part in np.array_split(ev, 3):
            dec.feed(part, 0.5)
        sym = dec.symbols()
        want = z["sym%d" % ci]
        assert np.array_equal(np.stack([sym["type"], sym["val"]], 1).astype(np.int32).reshape(-1, 2), want), ci
        fr, bits = dec.frames()
        assert fr["type"].tolist() == z["ftype%d" % ci].tolist(), ci
        assert fr["nbits"].tolist() == z["flen%d" % ci].tolist(), ci
        flat = np.concatenate(bits) if bits else np.zeros(0, np.uint8)
        assert np.array_equal(flat, z["fbits%d" % ci]), ci


def test_decoder_direction_switches():
    z = H.load_case("decoder_kat")
    evs = z["ev0"]
    ev = np.zeros(len(evs), oracle.EVENT_DTYPE)
    ev["v"], ev["d"], ev["type"] = evs[:, 0], evs[:, 1], evs[:, 2]
    both = oracle.Decoders(True, True)
    both.feed(ev, 0.5)
    for reader, tag, keep in ((True, False, 1), (False, True, 0)):
        one = oracle.Decoders(reader, tag)
        one.feed(ev, 0.5)
        want = both.symbols()
        want = want[want["type"] == keep]
        got = one.symbols()
        assert np.array_equal(got["val"], want["val"])


def test_encoders_known_answers():
    for rec in H.load_json("encoders.json"):
        for name, fn in (("miller", oracle.miller_encode), ("manchester", oracle.manchester_encode)):
            lv, du = fn(rec["bits"])
            want = rec[name]
            assert lv.tolist() == [p[0] for p in want]
            assert du.tolist() == [p[1] for p in want]


def test_miller_report_example():
    rec = H.load_json("miller_report_example.json")
    ev = np.zeros(len(rec["pulses"]), oracle.EVENT_DTYPE)
    for i, (v, d) in enumerate(rec["pulses"]):
        ev[i] = (i, d, v, 1, 0)
    dec = oracle.Decoders(True, True)
    dec.feed(ev, 1.0)
    assert dec.symbols()["val"].tolist() == rec["symbols"]
    assert rec["symbols"][:4] == [0, 1, 0, 1]  # report/report.pdf p.5


def test_fsm_tail_known_answers():
    for rec in H.load_json("fsm_tail.json"):
        fixed, flag = oracle.fix_ending(rec["bits"], rec["type"])
        assert fixed.tolist() == rec["fixed"]
        assert {0: "", 1: "EXTRA ERROR", 2: "MANY MORE ERROR"}[flag] == rec["msg"]
        par = oracle.check_parity(fixed)
        assert (None if par is None else par.tolist()) == rec["parity"]
        by, fl = oracle.print_enc(fixed)
        assert [[int(b), int(f)] for b, f in zip(by, fl)] == rec["enc"]


def test_crc_a_known_answers():
    """CRC_A against the reference's utilities.CRC (fixture: oracle/gen_golden_crc.py)."""
    cases = H.load_json("crc_a.json")
    assert cases[0]["data"] == [0x50, 0x00] and cases[0]["crc"] == [0x57, 0xCD]  # HLTA, the one every trace shows
    for rec in cases:
        assert oracle.crc_a(rec["data"]) == rec["crc"]
        assert oracle.check_crc(rec["data"] + rec["crc"]) == rec["check_good"] is True
        assert oracle.check_crc(rec["bad"]) == rec["check_bad"]


def test_empty_and_tiny_inputs():
    ts = oracle.TransitionSink(2e6)
    used, ev = ts.work(np.zeros(0, np.float32))
    assert used == 0 and ev is None
    used, ev = ts.work(np.ones(1999, np.float32))
    assert used == 1999 and ev is None
    used, ev = ts.work(np.ones(10, np.float32))  # warm-up completes with 1 item; rest is re-offered
    assert used == 1 and ev is None
    used, ev = ts.work(np.zeros(0, np.float32))  # work_stable still calls back with an empty list
    assert used == 0 and ev is not None and len(ev) == 0
