"""Known answers at the cpp.append_bit boundary: the reference's hand-annotated on-air bits of its Ultralight session
(outputs/ultralight_bits.txt -> tests/golden/ultralight_bits.json, oracle/gen_golden_bits.py).  Every byte is typed LSB
first with its odd parity bit under it; the hex value typed beside it disagrees once (a typo of the file)."""
import numpy as np
import pytest

from oracle import oracle
from usrp_nfc_b200 import synth

from . import helpers as H


def _frames():
    return H.load_json("ultralight_bits.json")["frames"]


def _typed_bits(fr):
    out = []
    for by in fr["bytes"]:
        out += by["bits"] + [by["parity"]]
    return out


def test_file_is_what_the_survey_says():
    fr = _frames()
    assert len(fr) == 18 and sum(len(f["bytes"]) for f in fr) == 132
    typos = [(by["line"], by["hex"]) for f in fr for by in f["bytes"] if by["typo"]]
    assert typos == [(383, 0x29)]  # the bits typed there are those of 0x49


def test_bytes_are_lsb_first_with_odd_parity():
    """utilities.py:52-63 as restated by synth.bytes_to_bits and checked by the oracle's fsm._check_parity (fsm.py:28-49)."""
    L = oracle.lib()
    for f in _frames():
        vals = [sum(b << i for i, b in enumerate(by["bits"])) for by in f["bytes"]]
        assert all(by["parity"] in (0, 1) for by in f["bytes"])
        assert synth.bytes_to_bits(vals) == _typed_bits(f)
        bits = np.array(_typed_bits(f), dtype=np.uint8)
        out = np.zeros(len(vals) + 1, dtype=np.uint8)
        n = L.nfc_check_parity(bits.ctypes.data, bits.size, out.ctypes.data)
        assert n == len(vals) and out[:n].tolist() == vals
        for by, v in zip(f["bytes"], vals):
            assert by["typo"] or by["hex"] == v


def test_frames_are_the_logged_session():
    """The same 18 frames, in order, as outputs/ultralight.out logs after its REQA (tests/golden/logged_frames.json)."""
    logged = H.load_json("logged_frames.json")["ultralight"][1:]
    fr = _frames()
    assert len(logged) == len(fr)
    for f, lg in zip(fr, logged):
        vals = [sum(b << i for i, b in enumerate(by["bits"])) for by in f["bytes"]]
        assert vals == lg["bytes"], lg["name"]


@pytest.mark.gpu
def test_device_frames_carry_the_annotated_bits():
    """The surrogate Ultralight capture through the CUDA path: frame k+1's bits are the typed bits of frame k of the file
    (reader frames followed by the end bit the Miller decoder hands over; a tag frame whose last typed bit is 1 arrives
    without it -- the Manchester decoder sees no edge after it -- and fsm._fix_ending, fsm.py:51-66, appends the start
    bit 1 again), and the device's frame tail (fsm._fix_ending / _check_parity) turns them back into the typed bytes."""
    from usrp_nfc_b200 import _cabi
    case = H.load_case("surrogate_ultralight")
    s = _cabi.Stream(2e6, hi_val=1.09)
    s.push_all(H.case_input(case))
    fr, bits = s.drain_frames_flat()
    s.close()
    typed = _frames()
    assert len(fr) == len(typed) + 1
    tails, by, flags = _cabi.frames_tail(fr, bits)
    for k, f in enumerate(typed):
        r = fr[k + 1]
        got = bits[int(r["bit_off"]): int(r["bit_off"]) + int(r["nbits"])].tolist()
        want = _typed_bits(f)
        if len(got) == len(want) - 1:
            assert int(r["type"]) == 0 and want[-1] == 1 and got == want[:-1], k
        else:
            assert got[: len(want)] == want, k
            assert len(got) - len(want) in (0, 1)
        t = tails[k + 1]
        vals = [sum(b << i for i, b in enumerate(x["bits"])) for x in f["bytes"]]
        assert int(t["parity_ok"]) == 1 and by[int(t["byte_off"]): int(t["byte_off"]) + int(t["nbytes"])].tolist() == vals
        assert not flags[int(t["byte_off"]): int(t["byte_off"]) + int(t["nbytes"])].any()
