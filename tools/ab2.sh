#!/bin/bash
# GPU check + A/B of the dense MINCO path (development aid): parity tests first, then throughput of the build with and
# without DFTPAV_NO_DENSE, then the alternative builds in gpurun_out_in/lib_<name>.so
mkdir -p gpurun_out
timeout 480 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
cp dftpav_b200/csrc/libdftpav_b200.so /tmp/lib_main.so
run() { timeout -k 10 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'value %.0f e2e %.0f evals/s %.3e parity %s'%(d['value'], d['e2e']['value'], d['evals_per_s'], d.get('parity_vs_oracle')))"; }
run main
DFTPAV_NO_DENSE=1 run nodense
for f in gpurun_out_in/lib_*.so; do n=$(basename $f .so); cp $f dftpav_b200/csrc/libdftpav_b200.so; run $n; done
cp /tmp/lib_main.so dftpav_b200/csrc/libdftpav_b200.so
timeout 120 python tools/phase_clock.py 2>&1 | tail -12
