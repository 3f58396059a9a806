#!/bin/bash
mkdir -p gpurun_out
for cl in 1 0; do
  echo "== OBS_CLUSTER=$cl"
  ORGYM_NET_OBS_CLUSTER=$cl INFO=0 python tools/net64_quick.py 2>&1 | grep "net64 step"
  ORGYM_NET_OBS_CLUSTER=$cl INFO=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"net_obs|net_jit_step" -s 8 -c 4 python tools/net64_quick.py 2>&1 | grep -E "net_obs_kernel|net_jit_step|duration|dram"
done
python -m pytest tests/test_netinv_gpu.py -m gpu -x -q -k "stream_jit or 64 or masked" 2>&1 | tail -2
