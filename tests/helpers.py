"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np

from usrp_nfc_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def case_input(case):
    """float32 envelope the reference's transition_sink saw for a stored capture."""
    return synth.envelope(synth.pcm_to_float(case["pcm"]))


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def summary():
    return load_json("SUMMARY.json")


def assert_events_equal(got, want, what="events"):
    assert len(got) == len(want), "%s: count %d != %d" % (what, len(got), len(want))
    for f in ("pos", "d", "v", "type"):
        if f in want.dtype.names and f in got.dtype.names:
            bad = np.nonzero(got[f] != want[f])[0]
            assert bad.size == 0, "%s: field %s first differs at %d: got %r want %r" % (
                what, f, bad[0], got[bad[0]], want[bad[0]])


def assert_symbols_equal(got, want, what="symbols"):
    assert len(got) == len(want), "%s: count %d != %d" % (what, len(got), len(want))
    for f in ("pos", "type", "val"):
        bad = np.nonzero(got[f] != want[f])[0]
        assert bad.size == 0, "%s: field %s first differs at %d" % (what, f, bad[0])


def assert_frames_equal(frames, frame_bits, case, what="frames"):
    assert len(frames) == len(case["fpos"]), "%s: count %d != %d" % (what, len(frames), len(case["fpos"]))
    assert np.array_equal(frames["pos"], case["fpos"]), what + ": closing positions differ"
    assert np.array_equal(frames["type"], case["ftype"]), what + ": types differ"
    assert np.array_equal(frames["nbits"], case["flen"]), what + ": lengths differ"
    flat = np.concatenate(frame_bits) if frame_bits else np.zeros(0, np.uint8)
    assert np.array_equal(flat, case["fbits"]), what + ": bits differ"
