#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
for t in 0 1; do
TG_FRAME_TAP=$t timeout 600 ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 4 -c 1 -f -o gpurun_out/r01_frame_pair_tap$t python bench.py --steps 1 --warmup 3 --clips 2 --frames 2 --no-e2e --no-cpu-baseline --no-train > gpurun_out/ncu_full_tap$t.log 2>&1
echo "ncu tap=$t rc=$?"
done
ls -la gpurun_out/*.ncu-rep
