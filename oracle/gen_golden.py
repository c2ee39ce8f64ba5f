"""Generate tests/golden/ and usrp_nfc_b200/data/sessions.json from the reference -- TEST INFRASTRUCTURE.

Runs ONLY in the build container (needs /root/reference).  It imports the reference's
own modules (oracle/refshim.py), drives them on seeded inputs and stores inputs +
reference outputs as small fixtures.  The GPU box has no /root/reference; tests there
use the committed fixtures.

    python oracle/gen_golden.py

Python 2 vs 3: the only arithmetic difference on this path is sum() (compensated since
3.12).  transition_sink.py:122 calls sum(ar); we bind a left-to-right `sum` into that
module's globals so the fixtures carry Python 2 results even for inexact window sums.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import refshim  # noqa: E402

refshim.install()
import packets  # noqa: E402  (must come first: packets -> fsm -> command -> packets cycle, packets.py:55)
import command  # noqa: E402
import fsm as ref_fsm  # noqa: E402
import manchester  # noqa: E402
import miller  # noqa: E402
import transition_sink  # noqa: E402

from usrp_nfc_b200 import synth  # noqa: E402


def _py2_sum(seq):
    s = 0
    for v in seq:
        s = s + v
    return s


transition_sink.sum = _py2_sum

GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "usrp_nfc_b200", "data")
OUTPUTS = os.path.join(os.path.dirname(refshim.REF_CODE), "outputs")


# ---------------------------------------------------------------- log parsing
def _name_types():
    m = {}
    for k, v in vars(command.CommandType).items():
        if isinstance(v, command.Command):
            m[v.name()] = v.packet_type()
    return m


def _hex_tokens(s):
    return re.findall(r"0[xX]([0-9A-Fa-f]{2})(!?)", s)


def norm_log(text):
    """Comparable form of a frame log: case-folded, blank lines and banners dropped, the
    spliced 'PROCESSING FINISHED' (outputs/1k_with_enc.out:1232) removed."""
    out = []
    for line in text.splitlines():
        line = line.replace("PROCESSING FINISHED", "")
        line = " ".join(line.split()).lower()
        if not line or line.startswith("linux;") or line.startswith("using volk"):
            continue
        if line.startswith("0x") and out and out[-1].startswith("0x"):
            out[-1] += " " + line  # raw line cut in two by the spliced text (outputs/1k_with_enc.out:1232-1233)
        else:
            out.append(line)
    return out


def parse_log(path):
    """-> list of frames {name, type, bytes, raw:[(byte, flagged)] or None}"""
    types = _name_types()
    frames, raw, cur = [], None, None
    with open(path) as f:
        lines = f.read().splitlines()
    for line in lines:
        line = line.replace("PROCESSING FINISHED", "").strip()
        if line.upper().startswith("COMMAND:"):
            cur = {"name": line.split(":", 1)[1].strip(), "bytes": [], "raw": raw}
            cur["type"] = types[cur["name"]]
            raw = None
            frames.append(cur)
        elif re.match(r"^(HEADER|EXTRA|CRC):", line, re.I):
            cur["bytes"].extend(int(h, 16) for h, _ in _hex_tokens(line))
        elif re.match(r"^0[xX]", line):
            raw = (raw or []) + [(int(h, 16), bang == "!") for h, bang in _hex_tokens(line)]
    return frames


def onair_bits(fr):
    """On-air bits of a logged frame (without start/end bits)."""
    if fr["raw"] is not None:
        bits = []
        for b, flagged in fr["raw"]:
            byte_bits = [(b >> i) & 1 for i in range(8)]
            ones = sum(byte_bits) & 1
            bits.extend(byte_bits)
            bits.append(ones if flagged else 1 - ones)  # fsm.py:124-127: '!' <=> parity bit == popcount&1
        return bits
    if fr["name"] in ("REQA", "WUPA"):
        return [(fr["bytes"][0] >> i) & 1 for i in range(7)]  # short frame, no parity
    return synth.bytes_to_bits(fr["bytes"], parity=True)


# ------------------------------------------------------------ reference runs
def run_reference(x, samp_rate, with_fsm, **kw):
    refshim.log.seek(0)
    refshim.log.truncate()
    ch = refshim.ReferenceChain(samp_rate, with_fsm=with_fsm, **kw).run(x)
    factor = 1e6 / samp_rate
    ev = np.zeros(len(ch.events), dtype=[("pos", "<i8"), ("d", "<i4"), ("v", "i1"), ("type", "i1")])
    for i, (pos, v, dur, t) in enumerate(ch.events):
        d = int(round(dur / factor))
        assert d * factor == dur, (d, factor, dur)
        ev[i] = (pos, d, v, t)
    sym = np.array([(p, t, b) for p, t, b in ch.symbols], dtype=[("pos", "<i8"), ("type", "i1"), ("val", "i1")])
    fpos = np.array([p for p, _, _ in ch.frames], dtype=np.int64)
    ftype = np.array([t for _, t, _ in ch.frames], dtype=np.int8)
    flen = np.array([len(b) for _, _, b in ch.frames], dtype=np.int32)
    fbits = np.array([b for _, _, bits in ch.frames for b in bits], dtype=np.uint8)
    st = ch.sink
    state = dict(ss=float(st._sum), cur_state=int(st._current_state), dur=int(st._dur), last_bit=int(st._last_bit),
                 index=int(st._index), ring=np.array(st._ar, dtype=np.float64))
    return dict(ev=ev, sym=sym, fpos=fpos, ftype=ftype, flen=flen, fbits=fbits, state=state,
                log=refshim.log.getvalue())


def save_case(name, pcm=None, x=None, **arrs):
    out = {}
    if pcm is not None:
        out["pcm"] = pcm
    if x is not None:
        out["x"] = x
    for k, v in arrs.items():
        if k == "state":
            for sk, sv in v.items():
                out["state_" + sk] = np.asarray(sv)
        elif k != "log":
            out[k] = v
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    os.makedirs(DATA, exist_ok=True)
    summary = {}

    # 1. session scripts from the reference's golden logs
    logs = {"classic1k": "1k_with_enc.out", "ultralight": "ultralight.out"}
    sessions, parsed = {}, {}
    for key, fn in logs.items():
        parsed[key] = parse_log(os.path.join(OUTPUTS, fn))
        sessions[key] = [(fr["type"], "".join(str(b) for b in onair_bits(fr))) for fr in parsed[key]]
    with open(os.path.join(DATA, "sessions.json"), "w") as f:
        json.dump(sessions, f, indent=0)
    expect = {k: [dict(name=fr["name"], type=fr["type"], bytes=fr["bytes"],
                       raw=None if fr["raw"] is None else [[b, int(fl)] for b, fl in fr["raw"]])
                  for fr in v] for k, v in parsed.items()}
    with open(os.path.join(GOLD, "logged_frames.json"), "w") as f:
        json.dump(expect, f)
    summary["sessions"] = {k: len(v) for k, v in sessions.items()}

    # 2. encoder known answers (reference encoders) and our generator's encoders must agree
    rng = np.random.default_rng(7)
    enc = []
    for n in [0, 1, 2, 7, 9, 18, 163]:
        for _ in range(3):
            bits = rng.integers(0, 2, n).tolist()
            mi = miller.miller_encoder.encode_bits(bits)
            ma = manchester.manchester_encoder.encode_bits(bits)
            assert mi == synth.miller_encode(bits) and ma == synth.manchester_encode(bits)
            enc.append(dict(bits=bits, miller=mi, manchester=ma))
    with open(os.path.join(GOLD, "encoders.json"), "w") as f:
        json.dump(enc, f)

    # 3. surrogate captures of the two logged sessions at 2 MS/s
    for key, seed in (("classic1k", 2000), ("ultralight", 1000)):
        frames = [(t, [int(c) for c in s]) for t, s in sessions[key]]
        pcm = synth.capture(frames, 2e6, seed, channel=synth.Channel(pause=0.01, tag_high=1.06))
        x = synth.envelope(synth.pcm_to_float(pcm))
        res = run_reference(x, 2e6, with_fsm=True, hi_val=1.09)
        got = norm_log(res["log"])
        with open(os.path.join(OUTPUTS, logs[key])) as f:
            want = norm_log(f.read())
        same = sum(a == b for a, b in zip(got, want))
        print("%s: %d samples, %d events, %d frames, log lines identical %d/%d (ours %d)" % (
            key, x.size, len(res["ev"]), len(res["fpos"]), same, len(want), len(got)))
        assert got == want, "surrogate capture does not reproduce the reference's golden log"
        save_case("surrogate_" + key, pcm=pcm, **res)
        summary["surrogate_" + key] = dict(samples=int(x.size), events=len(res["ev"]), frames=len(res["fpos"]),
                                           log_lines=len(want))

    # 4. other rates with pinned av_window / max_len; first frames of the ultralight session
    ul = [(t, [int(c) for c in s]) for t, s in sessions["ultralight"]]
    for rate, seed, nfr in ((13.56e6, 3000, 10), (20e6, 5000, 8)):
        p = synth.rate_params(rate)
        pcm = synth.capture(ul[:nfr], rate, seed, channel=synth.Channel(pause=0.03, tag_high=1.09, fade=0.05),
                            av_window=p["av_window"])
        x = synth.envelope(synth.pcm_to_float(pcm))
        res = run_reference(x, rate, with_fsm=False, hi_val=1.09, **p)
        name = "rate_%d" % int(rate / 1e4)
        print("%s: %d samples, %d events, %d frames" % (name, x.size, len(res["ev"]), len(res["fpos"])))
        save_case(name, pcm=pcm, **res)
        summary[name] = dict(samples=int(x.size), events=len(res["ev"]), frames=len(res["fpos"]), **p)

    # 5. slicer known answers on adversarial float streams (small windows, zeros, spikes, negatives)
    rng = np.random.default_rng(11)
    kat = {}
    cases = []
    for ci in range(24):
        L = int(rng.choice([1, 2, 5, 16, 64, 300]))
        mx = int(rng.choice([1, 3, 7, 50]))
        n = int(rng.integers(0, 3000))
        kind = ci % 6
        if kind == 0:
            x = rng.integers(0, 1 << 24, n).astype(np.float32) / np.float32(1 << 24)
        elif kind == 1:
            x = np.where(rng.random(n) < 0.3, 0.0, rng.random(n)).astype(np.float32)
        elif kind == 2:
            x = (0.25 * (1 + 0.02 * rng.standard_normal(n))).astype(np.float32)
            x[rng.random(n) < 0.05] = 1e-4
            x[rng.random(n) < 0.05] = 0.4
        elif kind == 3:
            x = np.zeros(n, dtype=np.float32)
            x[n // 2:] = rng.random(n - n // 2).astype(np.float32)
        elif kind == 4:
            x = (rng.standard_normal(n) * 10.0 ** rng.integers(-20, 20, n)).astype(np.float32)  # inexact sums, negatives
        else:
            base = np.repeat(rng.choice([0.25, 0.0004, 0.3], max(1, n // 9 + 1)), 9)[:n]
            x = (base * (1 + 0.01 * rng.standard_normal(n))).astype(np.float32)
        lo = float(rng.choice([0.1, 0.5]))
        hi = float(rng.choice([1.1, 1.05, 1.5]))
        chunks = [int(c) for c in rng.integers(1, 700, 5)]
        ch = refshim.ReferenceChain(2e6, with_fsm=False, lo_val=lo, hi_val=hi, av_window=L, max_len=mx)
        batches = ch.run_chunked(x, chunks)
        flat = [e for b in batches for e in b]
        st = ch.sink
        kat["x%d" % ci] = x
        kat["ev%d" % ci] = np.array([(v, int(round(d / 0.5)), t) for v, d, t in flat], dtype=np.int32).reshape(-1, 3)
        kat["nb%d" % ci] = np.array([len(b) for b in batches], dtype=np.int32)
        kat["ring%d" % ci] = np.array(st._ar, dtype=np.float64)
        cases.append(dict(L=L, mx=mx, lo=lo, hi=hi, chunks=chunks, ss=float(st._sum), cur_state=int(st._current_state),
                          dur=int(st._dur), last_bit=int(st._last_bit), index=int(st._index),
                          stable=bool(st.work == st.work_stable)))
    kat["meta"] = np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLD, "slicer_kat.npz"), **kat)
    summary["slicer_kat"] = len(cases)

    # 6. decoder + framing known answers on random event streams (alphabet superset v in -1..2)
    rng = np.random.default_rng(13)
    dk = {}
    for ci in range(40):
        n = int(rng.integers(1, 400))
        mx = 50
        if ci % 4 == 0:  # plausible traffic: encoder pulses with jitter, mixed with garbage
            evs = []
            for _ in range(6):
                bits = rng.integers(0, 2, int(rng.integers(1, 40))).tolist()
                if rng.random() < 0.5:
                    pulses, t = miller.miller_encoder.encode_bits(bits), 1
                    evs.append((1, mx, 1))
                    evs += [(lv, max(1, int(round(du * 2 + rng.normal(0, 0.6)))), t) for lv, du in pulses[:-1]]
                    evs.append((1, mx, 1))
                else:
                    pulses, t = manchester.manchester_encoder.encode_bits(bits), 0
                    evs.append((0, mx, 0))
                    evs += [(lv, max(1, int(round(du * 2 + rng.normal(0, 0.6)))), t) for lv, du in pulses]
                    evs.append((0, mx, 0))
            evs = [(v, min(d, mx), t) for v, d, t in evs]
        else:
            evs = [(int(rng.integers(-1, 3)), int(rng.integers(1, mx + 1)), int(rng.choice([-1, 0, 1, 1, 0])))
                   for _ in range(n)]
        cpp_syms, cpp_frames = [], []

        class Cpp(object):
            def __init__(self):
                self.pp = [packets.PacketProcessor(i) for i in range(2)]

            def append_bit(self, bit, t):
                cpp_syms.append((t, int(bit)))
                ret = self.pp[t].append_bit(bit)
                if ret:
                    cpp_frames.append((t, [int(b) for b in ret]))

        cpp = Cpp()
        rd, tg = miller.miller_decoder(cpp), manchester.manchester_decoder(cpp)
        for v, d, t in evs:
            if t == 0:
                tg.process_transition([(v, d * 0.5)])
            elif t == 1:
                rd.process_transition([(v, d * 0.5)])
        dk["ev%d" % ci] = np.array(evs, dtype=np.int32).reshape(-1, 3)
        dk["sym%d" % ci] = np.array(cpp_syms, dtype=np.int32).reshape(-1, 2)
        dk["ftype%d" % ci] = np.array([t for t, _ in cpp_frames], dtype=np.int32)
        dk["flen%d" % ci] = np.array([len(b) for _, b in cpp_frames], dtype=np.int32)
        dk["fbits%d" % ci] = np.array([b for _, bits in cpp_frames for b in bits], dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLD, "decoder_kat.npz"), **dk)
    summary["decoder_kat"] = 40

    # 7. the worked Miller example of report/report.pdf p.5
    out = []

    class Rec(object):
        def append_bit(self, bit, t):
            out.append(int(bit))

    rd = miller.miller_decoder(Rec())
    rd.process_transition([(0, 3), (1, 11), (0, 3), (1, 16), (0, 3), (1, 6)])
    assert out[:4] == [0, 1, 0, 1], out
    with open(os.path.join(GOLD, "miller_report_example.json"), "w") as f:
        json.dump(dict(pulses=[[0, 3], [1, 11], [0, 3], [1, 16], [0, 3], [1, 6]], symbols=out), f)

    # 8. host tail: _fix_ending / _check_parity / _print_enc (fsm.py:28-66,114-131)
    rng = np.random.default_rng(17)
    tail = []
    f = ref_fsm.fsm()
    for ci in range(60):
        n = int(rng.integers(1, 60))
        if ci % 3 == 0:
            bits = synth.bytes_to_bits(rng.integers(0, 256, n // 9 + 1).tolist())
            bits = bits[: len(bits) - int(rng.integers(0, 3))] + rng.integers(0, 2, int(rng.integers(0, 2))).tolist()
        else:
            bits = rng.integers(0, 2, n).tolist()
        t = int(rng.integers(0, 2))
        refshim.log.seek(0)
        refshim.log.truncate()
        fixed = f._fix_ending(list(bits), t)
        fix_msg = refshim.log.getvalue().strip()
        par = f._check_parity(list(fixed))
        refshim.log.seek(0)
        refshim.log.truncate()
        f._print_enc(list(fixed))
        enc = [[int(h, 16), int(b == "!")] for h, b in _hex_tokens(refshim.log.getvalue())]
        tail.append(dict(bits=bits, type=t, fixed=[int(b) for b in fixed], msg=fix_msg,
                         parity=None if par is None else [int(b) for b in par], enc=enc))
    with open(os.path.join(GOLD, "fsm_tail.json"), "w") as fo:
        json.dump(tail, fo)

    with open(os.path.join(GOLD, "SUMMARY.json"), "w") as fo:
        json.dump(summary, fo, indent=1)
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
