#!/bin/bash
# run quick parity with an alternative library build (development aid)
cp dftpav_b200/csrc/libdftpav_b200.so /tmp/lib_main.so
echo "== main build"; timeout 100 python tools/quick_gpu.py 3 64 8 2>&1 | tail -4
cp gpurun_out_in/lib_nopure.so dftpav_b200/csrc/libdftpav_b200.so
echo "== volatile-lds build"; timeout 100 python tools/quick_gpu.py 3 64 8 2>&1 | tail -4
cp /tmp/lib_main.so dftpav_b200/csrc/libdftpav_b200.so
