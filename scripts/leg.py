"""One `configs` leg of bench.py alone: python scripts/leg.py <samp_rate> <samples> [fade]   (e.g. 20e6 4e9)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rate, n = float(sys.argv[1]), float(sys.argv[2])
fade = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
sys.argv = ["bench.py"]
import bench  # noqa: E402
import torch  # noqa: E402
from usrp_nfc_b200 import _cabi  # noqa: E402


class A:
    tag_high = 1.07


A.fade = fade
peak, _ = bench.measured_peak_gbs()
r = bench.leg_stream(torch, _cabi, rate, n, 0, A, peak)
print({k: (float(v) if hasattr(v, "item") else v) for k, v in r.items() if k not in ("workload", "clocks")})
