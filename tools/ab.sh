#!/bin/bash
# A/B bench of alternative library builds placed in gpurun_out_in/lib_<name>.so (development aid)
cp dftpav_b200/csrc/libdftpav_b200.so /tmp/lib_main.so
run() { timeout -k 10 120 python bench.py --steps 4 --warmup 3 --no-cpu-baseline | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'value %.0f e2e %.0f evals/s %.3e'%(d['value'], d['e2e']['value'], d['evals_per_s']))"; }
run main
for f in gpurun_out_in/lib_*.so; do n=$(basename $f .so); cp $f dftpav_b200/csrc/libdftpav_b200.so; run $n; done
cp /tmp/lib_main.so dftpav_b200/csrc/libdftpav_b200.so
