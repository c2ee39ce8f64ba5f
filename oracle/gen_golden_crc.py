#!/usr/bin/env python
"""CRC_A known answers from the UNMODIFIED reference (code/utilities.py:26-46) -> tests/golden/crc_a.json.

TEST INFRASTRUCTURE, run in the build container only (needs /root/reference).  Separate from gen_golden.py so
that regenerating it leaves the other fixtures untouched.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refshim  # noqa: E402


def main():
    refshim.install()
    import utilities

    rng = np.random.default_rng(23)
    cases = []
    # the frames every ISO 14443A trace holds, then random payloads of every length the path sees
    fixed = [[0x50, 0x00], [0x93, 0x70, 0x88, 0x04, 0x4E, 0x5D, 0x9F], [0x30, 0x04], [0x60, 0x00], [0xE0, 0x80], []]
    for data in fixed + [rng.integers(0, 256, int(n)).tolist() for n in list(range(1, 21)) * 3]:
        data = [int(b) for b in data]
        crc = utilities.CRC.calculate_crc(list(data))
        good = data + [int(crc[0]), int(crc[1])]
        bad = list(good)
        if len(bad) > 2:
            bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        cases.append(dict(data=data, crc=[int(crc[0]), int(crc[1])], check_good=bool(utilities.CRC.check_crc(list(good))),
                          bad=bad, check_bad=bool(utilities.CRC.check_crc(list(bad)))))
    with open(os.path.join(os.path.dirname(HERE), "tests", "golden", "crc_a.json"), "w") as f:
        json.dump(cases, f)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
