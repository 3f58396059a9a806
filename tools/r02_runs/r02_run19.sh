#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_net64_tile.log
for v in tile32 tile64; do
  echo "== variant $v" >> $L
  ORGYM_B200_LIB=$PWD/or-gym-inventory_b200/csrc/variants/$v.so INFO=0 python tools/net64_quick.py 2>&1 | grep -E "step|rollout" >> $L
  ORGYM_B200_LIB=$PWD/or-gym-inventory_b200/csrc/variants/$v.so python tools/bench_quick.py net 2>&1 | grep -E "net step|net rollout" >> $L
done
echo "== default (tile 128)" >> $L
INFO=0 python tools/net64_quick.py 2>&1 | grep -E "step|rollout" >> $L
python tools/bench_quick.py net 2>&1 | grep -E "net step|net rollout" >> $L
cat $L
ORGYM_B200_LIB=$PWD/or-gym-inventory_b200/csrc/variants/tile32.so python -m pytest tests/test_netinv_gpu.py tests/test_canary_gpu.py -m gpu -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_net64_launches_t32.csv env ORGYM_B200_LIB=$PWD/or-gym-inventory_b200/csrc/variants/tile32.so python tools/prof_net64.py > gpurun_out/ncu.log 2>&1
grep -E "net_jit_step|net_obs" gpurun_out/r02_net64_launches_t32.csv | tail -2 | awk -F'","' '{print substr($5,1,20), $NF}'
