"""The drop-in blocks: same names, constructor parameters and callback contract as the reference."""
import inspect

import numpy as np
import pytest

from tests import helpers as H
from usrp_nfc_b200 import synth

pytestmark = pytest.mark.gpu


def test_transition_sink_block_contract():
    from usrp_nfc_b200.transition_sink import transition_sink
    sig = inspect.signature(transition_sink.__init__)
    names = list(sig.parameters)[1:7]
    assert names == ["samp_rate", "callback", "lo_val", "hi_val", "av_window", "max_len"]  # transition_sink.py:12
    assert [sig.parameters[n].default for n in names[2:]] == [0.1, 1.1, 2000, 50]
    case = H.load_case("surrogate_ultralight")
    x = H.case_input(case)
    batches = []
    blk = transition_sink(2e6, batches.append, hi_val=1.09)
    off, calls = 0, 0
    while off < x.size:
        used = blk.work([x[off: off + 8192]], None)
        off += used
        calls += 1
    flat = [e for b in batches for e in b]
    assert len(batches) == calls - 1  # warm-up call does not call back (transition_sink.py:109-125)
    assert len(flat) == len(case["ev"])
    for ((v, dur), t), w in zip(flat, case["ev"]):
        assert (v, t) == (int(w["v"]), int(w["type"])) and dur == int(w["d"]) * 0.5


def test_decoder_block_hands_frames_to_fsm():
    from usrp_nfc_b200.decoder import decoder
    assert list(inspect.signature(decoder.__init__).parameters)[1:8] == [
        "src", "dst", "repeat", "reader", "tag", "samp_rate", "emulator"]  # decoder.py:16
    case = H.load_case("surrogate_classic1k")
    got = []
    d = decoder(src=case["pcm"], samp_rate=2e6, on_frame=lambda bits, t: got.append((t, bits)))
    d.run()
    assert len(got) == len(case["fpos"])
    flat = np.array([b for _, bits in got for b in bits], dtype=np.uint8)
    assert np.array_equal(flat, case["fbits"])
    assert [t for t, _ in got] == case["ftype"].tolist()


def test_decoder_block_coalesces_small_work_calls():
    """coalesce: the items of successive work() calls go to the device together; the same frames in the same order."""
    from usrp_nfc_b200.decoder import decoder
    case = H.load_case("surrogate_classic1k")
    got = []
    d = decoder(src=case["pcm"], samp_rate=2e6, on_frame=lambda bits, t: got.append((t, bits)), coalesce=50000)
    assert d.run(chunk=8192) == len(case["pcm"])
    assert len(got) == len(case["fpos"]) and [t for t, _ in got] == case["ftype"].tolist()
    assert np.array_equal(np.array([b for _, bits in got for b in bits], dtype=np.uint8), case["fbits"])


def test_decoder_block_hands_frame_bytes_from_the_device_tail():
    """on_frame_bytes: what fsm.process_bits derives first (fsm.py:28-66,114-131; utilities.py:26-46), per frame, next to
    the bit lists of the same frames."""
    from oracle import oracle
    from usrp_nfc_b200.decoder import decoder
    case = H.load_case("surrogate_classic1k")
    frames, tails = [], []
    d = decoder(src=case["pcm"], samp_rate=2e6, on_frame=lambda bits, t: frames.append((t, bits)),
                on_frame_bytes=lambda by, fl, tl, t: tails.append((t, by, fl, tl.copy())))
    d.run()
    assert len(frames) == len(tails) == len(case["fpos"])
    for (t, bits), (t2, by, fl, tl) in zip(frames, tails):
        assert t == t2
        fixed, flag = oracle.fix_ending(bits, t)
        eb, ef = oracle.print_enc(fixed)
        assert by == eb.tolist() and fl == ef.tolist() and int(tl["fix_flag"]) == flag and int(tl["nbits"]) == fixed.size
        par = oracle.check_parity(fixed)
        assert bool(tl["parity_ok"]) == (par is not None and par.size > 0)
    # bytes only, no fsm module anywhere: the bit lists are never built
    only = []
    d2 = decoder(src=case["pcm"], samp_rate=2e6, on_frame_bytes=lambda by, fl, tl, t: only.append(by))
    d2.run()
    assert only == [by for _, by, _, _ in tails]


def test_decoder_block_with_fake_fsm_and_emulator():
    from usrp_nfc_b200.decoder import decoder

    class FakeFsm(object):
        class fsm(object):
            def __init__(self, callback=None):
                self.callback, self.frames = callback, []

            def process_bits(self, bits, t):
                self.frames.append((t, len(bits)))

            def process_outgoing(self, bits, cmd):
                return bits

    class Emu(object):
        def process_packet(self, cmd, struct):
            pass

        def set_encoder(self, enc):
            self.enc = enc

    case = H.load_case("surrogate_ultralight")
    emu = Emu()
    d = decoder(src=case["pcm"], samp_rate=2e6, emulator=emu, fsm=FakeFsm)
    d.run(chunk=5000)
    assert emu.enc == d._fsm.process_outgoing and d._fsm.callback == emu.process_packet  # packets.py:88-90
    assert [n for _, n in d._fsm.frames] == case["flen"].tolist()


def test_view_frames_is_the_drain_without_the_copy():
    import numpy as np
    from tests import helpers as H
    from usrp_nfc_b200 import _cabi
    case = H.load_case("rate_1356")
    x = H.case_input(case)
    kw = dict(hi_val=1.09, av_window=13560, max_len=339)
    a = _cabi.Stream(13.56e6, outputs=_cabi.OUT_FRAMES, **kw)
    b = _cabi.Stream(13.56e6, outputs=_cabi.OUT_FRAMES, **kw)
    a.push_all(x)
    b.push_all(x)
    fr, bits = a.drain_frames_flat()
    vf, b0, b1 = b.view_frames()
    assert len(vf) == len(fr) > 0
    assert np.array_equal(vf["pos"], fr["pos"]) and np.array_equal(vf["nbits"], fr["nbits"]) and np.array_equal(vf["type"], fr["type"])
    for f, g in zip(fr, vf):
        src = b0 if g["type"] == 0 else b1
        assert np.array_equal(bits[f["bit_off"]: f["bit_off"] + f["nbits"]], src[g["bit_off"]: g["bit_off"] + g["nbits"]])
    b.release_frames()
    assert len(b.view_frames()[0]) == 0
