#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_newsvendor_gpu.py -m gpu -x -q 2>&1 | tail -2
python tools/bench_quick.py nv
