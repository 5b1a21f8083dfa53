#!/bin/bash
# L1 prefetch hints (prefetch.global.L1) for the two-loop history: microbench, lone phases, throughput A/B
for d in 0 2 3 4; do echo "== prefetch distance $d"; timeout 60 gpurun_out_in/tlb_pf$d 2>&1 | grep "product"; done | tee gpurun_out/two_loop_bench_r2t.txt
for f in gpurun_out_in/fine_pf*.so; do echo "== $f"; timeout 120 python tools/lone_phases.py $f 2>&1 | grep -v warning | grep "kernel\|two-loop\|sweep"; done | tee gpurun_out/lone_phases_r2t.txt
bash tools/ab.sh 2>&1 | tee gpurun_out/ab_r2t.txt
