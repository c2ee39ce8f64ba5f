#!/bin/bash
# A/B of two builds of the library on ONE box: the bench step alternately with build_variants/libold.so and the in-tree build.
# The other build (build_variants/ is not tracked):  mkdir -p build_variants/old && git archive <commit> usrp_nfc_b200/csrc include |
#   tar -x -C build_variants/old && make -C build_variants/old/usrp_nfc_b200/csrc OUT=$PWD/build_variants/libold.so
mkdir -p gpurun_out
for r in 1 2; do
for v in old new; do
  if [ $v = old ]; then export USRP_NFC_B200_LIB=$PWD/build_variants/libold.so; else unset USRP_NFC_B200_LIB; fi
  timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-selfcheck --e2e-samples 1e6 > gpurun_out/ab_${v}_$r.json 2> gpurun_out/ab_${v}_$r.err
  python - <<PY
import json
b=json.load(open("gpurun_out/ab_${v}_$r.json"))
print("$v $r: ms %.2f dev %.2f slicer %.2f"%(b["ms_per_step"],b["device_ms_per_step"],b["slicer_ms_per_step"]))
PY
done; done
