#!/bin/bash
mkdir -p gpurun_out
timeout 1200 bash tools/ab.sh > gpurun_out/ab_r2j.log 2>&1; cat gpurun_out/ab_r2j.log
