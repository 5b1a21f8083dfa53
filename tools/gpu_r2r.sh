#!/bin/bash
tag=r2r
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$tag.log 2>&1; echo "ncu list exit $?"
bash tools/prof_solve.sh $tag
bash tools/prof_configs.sh $tag
