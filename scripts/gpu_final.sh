#!/bin/bash
# end-of-round artefacts: bench line + launch list + ncu --set full of the streaming slicer, then compute-sanitizer
bash scripts/gpu_art.sh
( timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/t_sanitize.py 2>&1 | grep -E "^kind|RACECHECK SUMMARY|hazard" | head -20
  timeout 500 compute-sanitizer --tool memcheck python scripts/t_sanitize.py 2>&1 | grep -E "^kind|ERROR SUMMARY|Invalid" | head -20 ) > gpurun_out/sanitizer.txt 2>&1
cat gpurun_out/sanitizer.txt
