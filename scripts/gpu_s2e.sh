#!/bin/bash
# small-window streaming kernel: full GPU suite, then the 2 MS/s leg against the first-generation kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -25 gpurun_out/pytest_gpu.log
python - <<'PY'
import json, time, sys
sys.argv=["bench.py"]
import bench, torch
from usrp_nfc_b200 import _cabi
class A: tag_high=1.07; fade=0.05
peak,_=bench.measured_peak_gbs()
for rate,n in ((2e6,2e9),(13.56e6,4e9)):
    r=bench.leg_stream(torch,_cabi,rate,n,0,A,peak)
    print({k:(float(v) if hasattr(v,"item") else v) for k,v in r.items() if k not in ("workload","clocks")})
PY
