#!/bin/bash
# lone-problem phase split with the fine L-BFGS counters (diagnostic build)
cd dftpav_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DDFTPAV_PHASE_FINE=1 \
  -shared -o /tmp/libdftpav_fine.so dftpav_kernels.cu dftpav_api.cu dftpav_multi.cu -ldl && cd ../..
python tools/lone_phases.py /tmp/libdftpav_fine.so 2>&1 | tee gpurun_out/lone_phases_r2n.txt
