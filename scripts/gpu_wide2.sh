#!/bin/bash
# GPU box: wide vs tall frame kernel at 4/8 clips per launch + per-segment trace at 2/4 clips.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
for c in 4 8; do
  timeout 600 python bench.py --steps 3 --warmup 3 --clips $c --no-train --no-cpu-baseline --no-e2e > gpurun_out/bench_wide_c$c.log 2>&1; echo "bench wide clips=$c rc=$?"; tail -1 gpurun_out/bench_wide_c$c.log | cut -c1-120;  grep -o '"clocks": {[^}]*}' gpurun_out/bench_wide_c$c.log
  TG_FRAME_WIDE=0 timeout 600 python bench.py --steps 3 --warmup 3 --clips $c --no-train --no-cpu-baseline --no-e2e > gpurun_out/bench_tall_c$c.log 2>&1; echo "bench tall clips=$c rc=$?"; tail -1 gpurun_out/bench_tall_c$c.log | cut -c1-120; grep -o '"clocks": {[^}]*}' gpurun_out/bench_tall_c$c.log
done
TG_N=4 timeout 300 python scripts/frame_trace.py > gpurun_out/frame_trace_wide_n4.txt 2>&1; echo "trace rc=$?"; head -1 gpurun_out/frame_trace_wide_n4.txt
TG_N=2 timeout 300 python scripts/frame_trace.py > gpurun_out/frame_trace_wide_n2.txt 2>&1; echo "trace rc=$?"; head -1 gpurun_out/frame_trace_wide_n2.txt
