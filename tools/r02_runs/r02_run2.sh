#!/bin/bash
# round 2, GPU call 2: network tests through the new observation kernel, then net64 timings old/new + launch list
mkdir -p gpurun_out
python -m pytest tests/test_netinv_gpu.py tests/test_canary_gpu.py tests/test_invmgmt_gpu.py -m gpu -x -q > gpurun_out/r02_nettests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_nettests.log
tail -5 gpurun_out/r02_nettests.log
L=gpurun_out/r02_net64.log
for info in 0 1; do
echo "== INFO=$info fused obs pass (round-1 form)" >> $L; ORGYM_NET_OBS_TMA=0 INFO=$info python tools/net64_quick.py >> $L 2>&1
echo "== INFO=$info TMA-staged obs kernel" >> $L; ORGYM_NET_OBS_TMA=1 INFO=$info python tools/net64_quick.py >> $L 2>&1
done
for mb in 3 5 6; do echo "== INFO=0 TMA obs, MINBLOCKS_STEP=$mb" >> $L; ORGYM_NET_JIT_MINBLOCKS_STEP=$mb INFO=0 python tools/net64_quick.py >> $L 2>&1; done
for th in 64 256; do echo "== INFO=0 TMA obs, THREADS=$th" >> $L; ORGYM_NET_JIT_THREADS=$th INFO=0 python tools/net64_quick.py >> $L 2>&1; done
cat $L
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_net64_launches.csv python tools/prof_net64.py > gpurun_out/ncu.log 2>&1
grep -E "net_jit_step|net_obs" gpurun_out/r02_net64_launches.csv | tail -6
echo "== inv sweep (magic conversion)" ; ORGYM_JIT_CACHE=0 python tools/bench_quick.py inv
