#!/bin/bash
# 2-GPU session: NCCL paths (dftpav_solve_multi, torchrun sharding with the real solver) and the 2-rank bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_multi.py -q -s > gpurun_out/multi_2gpu.log 2>&1; echo "multi rc $?"; tail -5 gpurun_out/multi_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc $?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_2gpu.json') if l.startswith('{')][-1]); print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['per_rank'], d['convergence_flag']['all_ranks_all_accepted'])"
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1of2.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_1of2.json')); print('N=1 value', d['value'], 'e2e', d['e2e']['value'])"
