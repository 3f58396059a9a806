#!/bin/bash
# round 2, GPU call 1: full gpu test suite, then timing sweep of the re-specialised serial rollout kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_gputests.log
tail -5 gpurun_out/r02_gputests.log
for mb in 4 6 8 10 12; do
  echo "== ORGYM_INV_JIT_MINBLOCKS=$mb" >> gpurun_out/r02_inv_sweep.log
  ORGYM_JIT_CACHE=0 ORGYM_INV_JIT_MINBLOCKS=$mb python tools/bench_quick.py inv >> gpurun_out/r02_inv_sweep.log 2>&1
done
echo "== ORGYM_INV_JIT_INT=0 (fma chain)" >> gpurun_out/r02_inv_sweep.log
ORGYM_JIT_CACHE=0 ORGYM_INV_JIT_INT=0 python tools/bench_quick.py inv >> gpurun_out/r02_inv_sweep.log 2>&1
cat gpurun_out/r02_inv_sweep.log
