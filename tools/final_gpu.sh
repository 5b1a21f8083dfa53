#!/bin/bash
# Round-end validation + evidence capture on the GPU box (everything lands in gpurun_out/)
tag=${1:-r1h}
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_$tag.log
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout -k 10 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 1800 gpurun_out/bench_$tag.json
timeout -k 10 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; tail -c 600 gpurun_out/bench_ref_$tag.json
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$tag.log 2>&1; echo "ncu list exit $?"
bash tools/prof_solve.sh $tag
timeout -k 10 300 python tools/all_configs.py > gpurun_out/all_configs_$tag.log 2>&1; cp gpurun_out/all_configs_r1.json gpurun_out/all_configs_$tag.json; tail -4 gpurun_out/all_configs_$tag.log
echo "latency, one problem per SM (B=148): team 1 then team 4"
timeout -k 10 60 python tools/quick_gpu.py 3 148 0 | grep "solve kernel" | tail -1
cp dftpav_b200/csrc/libdftpav_b200.so /tmp/lib_main.so; cp gpurun_out_in/lib_t4.so dftpav_b200/csrc/libdftpav_b200.so
timeout -k 10 60 python tools/quick_gpu.py 3 148 0 | grep "solve kernel" | tail -1
timeout -k 10 60 python tools/quick_gpu.py 3 1 0 | grep "solve kernel" | tail -1
cp /tmp/lib_main.so dftpav_b200/csrc/libdftpav_b200.so
timeout -k 10 60 python tools/quick_gpu.py 3 1 0 | grep "solve kernel" | tail -1
