#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=300 -x > gpurun_out/t_gen_pair.log 2>&1; echo "gen tests (pair+tap) rc=$?"; tail -3 gpurun_out/t_gen_pair.log
TG_FRAME_TAP=0 timeout 600 python -m pytest tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=300 -x > gpurun_out/t_gen_pair_wide.log 2>&1; echo "gen tests (pair, N-stacked) rc=$?"; tail -3 gpurun_out/t_gen_pair_wide.log
sel='total|res8.0|ct2.0|ct3.2 128->128 c0|convT128 c0|ct6|out 64|^stat'
for cfg in "1 1" "1 0" "0 0"; do
  set -- $cfg
  TG_FRAME_PAIR=$1 TG_FRAME_TAP=$2 TG_N=2 timeout 120 python scripts/frame_trace.py > gpurun_out/trace_pair$1_tap$2.txt 2>&1
  echo "== pair=$1 tap=$2"; grep -E "$sel" gpurun_out/trace_pair$1_tap$2.txt | cut -c1-22,60-130
done
for cfg in "1 1" "1 0"; do
set -- $cfg
TG_FRAME_PAIR=$1 TG_FRAME_TAP=$2 timeout 300 python bench.py --steps 3 --warmup 3 --clips 2 --no-train --no-cpu-baseline --no-e2e > gpurun_out/bench_pair$1_tap$2.log 2>&1; echo "bench pair=$1 tap=$2 rc=$?"; tail -1 gpurun_out/bench_pair$1_tap$2.log | cut -c1-140; grep -o '"clocks": {[^}]*}' gpurun_out/bench_pair$1_tap$2.log
done
