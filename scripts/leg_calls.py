"""The work()-sized-calls leg of bench.py alone (configs.work_calls_2MS), twice: python scripts/leg_calls.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ["bench.py"]
import bench  # noqa: E402
import torch  # noqa: E402
from usrp_nfc_b200 import _cabi  # noqa: E402


class A:
    tag_high = 1.07
    fade = 0.05


for rep in range(2):
    r = bench.leg_work_calls(torch, _cabi, 0, A)
    print(os.environ.get("USRP_NFC_B200_LIB", "in-tree"), rep, "per call %.3f ms (%.1f Msamples/s), coalesced %.1f Msamples/s" % (
        r["per_call"]["ms_per_call"], r["per_call"]["value"], r["coalesce_262144"]["value"]))
