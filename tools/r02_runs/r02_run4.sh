#!/bin/bash
# round 2, GPU call 4: report kernels + obs v3 tests; net64 timing with prefetch knob; newsvendor min-blocks variants
mkdir -p gpurun_out
python -m pytest tests/test_invmgmt_gpu.py tests/test_netinv_gpu.py -m gpu -x -q > gpurun_out/r02_tests4.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_tests4.log
tail -15 gpurun_out/r02_tests4.log
L=gpurun_out/r02_net64_v3.log
for pf in 0 1 2 3; do echo "== INFO=0 obs v3 PREFETCH=$pf" >> $L; ORGYM_NET_JIT_PREFETCH=$pf INFO=0 python tools/net64_quick.py 2>&1 | grep step >> $L; done
echo "== INFO=0 obs v3 PREFETCH=1 THREADS=256" >> $L; ORGYM_NET_JIT_THREADS=256 INFO=0 python tools/net64_quick.py 2>&1 | grep step >> $L
echo "== INFO=0 obs v3 PREFETCH=2 THREADS=256" >> $L; ORGYM_NET_JIT_PREFETCH=2 ORGYM_NET_JIT_THREADS=256 INFO=0 python tools/net64_quick.py 2>&1 | grep step >> $L
echo "== INFO=1 obs v3 PREFETCH=1" >> $L; INFO=1 python tools/net64_quick.py 2>&1 | grep step >> $L
cat $L
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_net64_launches_v3.csv python tools/prof_net64.py > gpurun_out/ncu.log 2>&1
grep -E "net_jit_step|net_obs" gpurun_out/r02_net64_launches_v3.csv | tail -4 | cut -d, -f5,15
for v in nv4 nv5; do echo "== newsvendor variant $v"; ORGYM_B200_LIB=$PWD/or-gym-inventory_b200/csrc/variants/$v.so python tools/bench_quick.py nv 2>&1 | grep rollout; done | tee gpurun_out/r02_nv_variants.log
