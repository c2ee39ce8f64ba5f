#!/bin/bash
mkdir -p gpurun_out
( timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/t_sanitize.py 2>&1 | grep -E "^kind|^corner|^small|RACECHECK SUMMARY|hazard|Read access|Write access" | head -40
  timeout 900 compute-sanitizer --tool memcheck python scripts/t_sanitize.py 2>&1 | grep -E "^kind|^corner|^small|ERROR SUMMARY|Invalid" | head -20 ) > gpurun_out/sanitizer.txt 2>&1
cut -c1-250 gpurun_out/sanitizer.txt
