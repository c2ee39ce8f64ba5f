#!/bin/bash
# A/B of prebuilt library variants: bash scripts/gpu_ab.sh v0 v1 ...   (build_variants/<name>.so)
for v in "$@"; do
  cp build_variants/$v.so usrp_nfc_b200/libusrp_nfc_b200.so
  echo "#### $v"
  bash scripts/gpu_b.sh "--samples 1e10 --steps 3" | tail -1 | sed 's/.*slicer_ms/slicer_ms/'
done
