#!/usr/bin/env python
"""The reference's hand-annotated on-air bits of the Ultralight session (outputs/ultralight_bits.txt) ->
tests/golden/ultralight_bits.json: per frame the bytes as typed (LSB first), the hex value typed beside them and the
parity bit typed under them.

TEST INFRASTRUCTURE, run in the build container only (needs /root/reference).  The file is free text: a frame is the run
of byte lines ("1100 1001 93") under a heading; a line holding a single bit after a byte line is that byte's parity bit,
single bits elsewhere are start / end bits.  The known typo of the file (line 383: hex 29 typed beside the bits of 49,
SURVEY.md section 4) is kept as typed and flagged.
"""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/outputs/ultralight_bits.txt"
BYTE = re.compile(r"^([01]{4}) ([01]{4}) ([0-9A-Fa-f]{2})\b")
BIT = re.compile(r"^([01])\s*(<-.*)?$")


def main():
    frames, cur, last_was_byte = [], None, False
    with open(SRC) as f:
        lines = f.read().split("\n")
    for no, raw in enumerate(lines, 1):
        line = raw.strip()
        m = BYTE.match(line)
        if m:
            if cur is None:
                cur = dict(line=no, bytes=[])
            bits = [int(c) for c in m.group(1) + m.group(2)]
            cur["bytes"].append(dict(bits=bits, hex=int(m.group(3), 16), parity=None, line=no))
            last_was_byte = True
            continue
        b = BIT.match(line)
        if b and cur is not None and last_was_byte:
            cur["bytes"][-1]["parity"] = int(b.group(1))
            last_was_byte = False
            continue
        last_was_byte = False
        if not b and line and cur is not None and cur["bytes"]:  # a heading or a remark ends the frame
            frames.append(cur)
            cur = None
    if cur is not None and cur["bytes"]:
        frames.append(cur)
    for fr in frames:
        for by in fr["bytes"]:
            val = sum(bit << i for i, bit in enumerate(by["bits"]))  # LSB first
            by["typo"] = val != by["hex"]
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "ultralight_bits.json")
    with open(out, "w") as f:
        json.dump(dict(source="outputs/ultralight_bits.txt", frames=frames), f)
    nb = sum(len(fr["bytes"]) for fr in frames)
    print(len(frames), "frames,", nb, "bytes,", sum(by["typo"] for fr in frames for by in fr["bytes"]), "typed hex values that disagree with their bits")


if __name__ == "__main__":
    main()
