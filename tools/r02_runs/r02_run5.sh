#!/bin/bash
# round 2, GPU call 5: report tests; net64 sync/threads sweep; new bench.py smoke (short)
mkdir -p gpurun_out
python -m pytest tests/test_invmgmt_gpu.py -m gpu -x -q -k "report or evaluate" > gpurun_out/r02_tests5.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_tests5.log
tail -5 gpurun_out/r02_tests5.log
L=gpurun_out/r02_net64_sync.log
for cfg in "0 128" "1 128" "0 256" "1 256" "0 512" "1 512"; do set -- $cfg
  echo "== INFO=0 PREFETCH=0 SYNC=$1 THREADS=$2" >> $L
  ORGYM_NET_JIT_PREFETCH=0 ORGYM_NET_JIT_SYNC=$1 ORGYM_NET_JIT_THREADS=$2 INFO=0 python tools/net64_quick.py 2>&1 | grep step >> $L
done
echo "== GROUP=8 SYNC=1 THREADS=256" >> $L; ORGYM_NET_JIT_PREFETCH=0 ORGYM_NET_JIT_GROUP=8 ORGYM_NET_JIT_SYNC=1 ORGYM_NET_JIT_THREADS=256 INFO=0 python tools/net64_quick.py 2>&1 | grep step >> $L
echo "== GROUP=2 SYNC=1 THREADS=256" >> $L; ORGYM_NET_JIT_PREFETCH=0 ORGYM_NET_JIT_GROUP=2 ORGYM_NET_JIT_SYNC=1 ORGYM_NET_JIT_THREADS=256 INFO=0 python tools/net64_quick.py 2>&1 | grep step >> $L
cat $L
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/r02_bench_smoke.json 2> gpurun_out/r02_bench_smoke.err; echo "bench exit $?"
tail -3 gpurun_out/r02_bench_smoke.err; head -c 3000 gpurun_out/r02_bench_smoke.json
timeout 300 python bench.py --workload netinv64_mlp --steps 4 --warmup 3 --no-extras > gpurun_out/r02_bench_mlp.json 2> gpurun_out/r02_bench_mlp.err; echo "bench mlp exit $?"
tail -3 gpurun_out/r02_bench_mlp.err; head -c 1500 gpurun_out/r02_bench_mlp.json
