#!/bin/bash
# compute-sanitizer over a small workload that touches every kernel (memcheck: out-of-bounds / misaligned accesses;
# synccheck: invalid barrier / __syncwarp use).  Output -> gpurun_out/sanitizer_$1_<tool>.txt
tag=${1:-x}
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/sanitizer_${tag}_$tool.txt 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|ok" gpurun_out/sanitizer_${tag}_$tool.txt | tail -8
done
