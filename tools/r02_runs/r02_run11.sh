#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_newsvendor_gpu.py tests/test_random_configs_gpu.py -m gpu -x -q -k "newsvendor or nv" > gpurun_out/r02_tests11.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_tests11.log
tail -4 gpurun_out/r02_tests11.log
python tools/bench_quick.py nv 2>&1 | grep rollout
python bench.py --workload newsvendor --no-extras --steps 5 2>/dev/null | head -c 600
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r02_nv_launches.csv python tools/prof_r02.py nv > /dev/null 2>&1
grep -E "nv_level|nv_rollout" gpurun_out/r02_nv_launches.csv | awk -F'","' '{print substr($5,1,30), $NF}'
