#!/bin/bash
# shortest possible A/B (development aid): one bench of the in-tree build, one of each gpurun_out_in/lib_*.so, then the
# small-batch latency of the last variant
run() { timeout -k 5 60 python bench.py --steps 3 --warmup 3 --no-cpu-baseline | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'value %.0f e2e %.0f'%(d['value'], d['e2e']['value']))"; }
cp dftpav_b200/csrc/libdftpav_b200.so /tmp/lib_main.so
run main
for f in gpurun_out_in/lib_*.so; do n=$(basename $f .so); cp $f dftpav_b200/csrc/libdftpav_b200.so; run $n; timeout -k 5 40 python tools/small_batches.py 2>/dev/null | head -4; done
cp /tmp/lib_main.so dftpav_b200/csrc/libdftpav_b200.so
