#!/bin/bash
# A/B on one box, interleaved: env assignments given as arguments, e.g. gpu_ab2.sh "TG_RGBX=1" "TG_RGBX=0"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
for rep in 1 2; do
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --clips 2 --no-train --no-glue --no-cpu-baseline --no-e2e > gpurun_out/bench_ab2.log 2>&1
  echo "$cfg: $(tail -1 gpurun_out/bench_ab2.log | grep -o '"value": [0-9.]*' | head -1) $(grep -o '"sm_mhz": [0-9]*' gpurun_out/bench_ab2.log) $(grep -o '"avg_launch_us": [0-9.]*' gpurun_out/bench_ab2.log)"
done
done
