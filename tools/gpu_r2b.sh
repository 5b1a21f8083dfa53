#!/bin/bash
# round-2 GPU session B: two-loop microbenchmark, tests (strict, parity, sampler), ncu captures
mkdir -p gpurun_out
timeout 300 ./gpurun_out_in/two_loop_bench > gpurun_out/two_loop_bench_r2b.txt 2>&1; cat gpurun_out/two_loop_bench_r2b.txt
timeout 900 python -m pytest tests/test_gpu_strict.py tests/test_gpu_sample.py -q -s > gpurun_out/strict_r2b.log 2>&1; echo "strict+sample rc $?"; tail -4 gpurun_out/strict_r2b.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -s > gpurun_out/parity_r2b.log 2>&1; echo "parity rc $?"; tail -4 gpurun_out/parity_r2b.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
# ncu: one full capture of the batch-4096 solve kernel, one of a lone long problem (source-level stalls of the two-loop)
timeout 420 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f -o gpurun_out/prof_solve_r2b python tools/quick_gpu.py 3 4096 0 > gpurun_out/ncu_solve_r2b.log 2>&1; echo "ncu solve rc $?"
timeout 420 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 2 -c 1 -f -o gpurun_out/prof_lone_r2b python tools/lone_problem.py 2 > gpurun_out/ncu_lone_r2b.log 2>&1; echo "ncu lone rc $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_r2b.log 2>&1; echo "ncu list rc $?"
ls -la gpurun_out/*r2b*
