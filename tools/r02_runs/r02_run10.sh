#!/bin/bash
mkdir -p gpurun_out
export ORGYM_NET_STREAM_AOT=1
python tools/prof_r02.py net64 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'net_stream_step_kernel' -s 6 -c 1 -o gpurun_out/r02_net64_aot \
      python tools/prof_r02.py net64 > gpurun_out/ncu_net64_aot.log 2>&1
ls -la gpurun_out/r02_net64_aot.ncu-rep
