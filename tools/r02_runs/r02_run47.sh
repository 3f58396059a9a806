#!/bin/bash
mkdir -p gpurun_out
python tools/prof_r02.py report || exit 1
ncu --set full --clock-control none --import-source on -k regex:"report_" -o gpurun_out/r02_report python tools/prof_r02.py report > gpurun_out/ncu_report.log 2>&1
ls -la gpurun_out/r02_report.ncu-rep
