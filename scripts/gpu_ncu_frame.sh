#!/bin/bash
# GPU box: one ncu --set full capture (with source-level sampling) of the frame kernel at 1 clip / launch.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 4 -c 1 -o gpurun_out/r01_frame_v3 -f python bench.py --steps 1 --warmup 3 --clips ${CLIPS:-1} --frames 2 --no-e2e --no-cpu-baseline --no-train > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
