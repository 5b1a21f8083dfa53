#!/bin/bash
# Round-end validation + evidence capture on the GPU box (everything lands in gpurun_out/)
tag=${1:-r2}
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -k 10 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 2500 gpurun_out/bench_$tag.json
timeout -k 10 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; tail -c 900 gpurun_out/bench_ref_$tag.json
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$tag.log 2>&1; echo "ncu list exit $?"
bash tools/prof_solve.sh $tag
timeout -k 10 400 python tools/all_configs.py > gpurun_out/all_configs_$tag.log 2>&1; cp gpurun_out/all_configs_latest.json gpurun_out/all_configs_$tag.json; tail -4 gpurun_out/all_configs_$tag.log | cut -c1-180
timeout -k 10 200 python tools/small_batches.py > gpurun_out/small_$tag.log 2>&1; cat gpurun_out/small_$tag.log
