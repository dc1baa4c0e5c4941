#!/bin/bash
# Same-box A/B of frame-kernel builds: arguments are "ENV=.. [ENV=..] [lib=<variant>]" groups, interleaved twice.
# Variant libraries are built here by the caller (python pytorch-tecogan_b200/build.py --variant NAME -D...).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1 || { tail -n 20 gpurun_out/build.log; exit 1; }
if [ -n "$TG_AB_TESTS" ]; then
  timeout 900 python -m pytest $TG_AB_TESTS -x -q -m gpu 2>&1 | tail -n 5
fi
for rep in 1 2; do
for cfg in "$@"; do
  lib=""; envs=""
  for tok in $cfg; do
    case $tok in lib=*) lib="--lib pytorch-tecogan_b200/libtecogan_b200.${tok#lib=}.so";; *) envs="$envs $tok";; esac
  done
  env $envs timeout 300 python bench.py --steps 5 --warmup 3 --clips 2 --no-train --no-glue --no-cpu-baseline --no-e2e --no-cfg3 --no-torch-gpu $lib > gpurun_out/bench_ab.log 2>&1
  echo "[$cfg]: $(tail -n 1 gpurun_out/bench_ab.log | grep -o '"value": [0-9.]*' | head -n 1) $(grep -o '"sm_mhz": [0-9]*' gpurun_out/bench_ab.log) $(grep -o '"avg_launch_us": [0-9.]*' gpurun_out/bench_ab.log | head -n 1) $(grep -o '"frac": [0-9.]*' gpurun_out/bench_ab.log | head -n 1)"
done
done
