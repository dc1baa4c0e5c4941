#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
TG_FRAME_TAP=${TAP:-0} timeout 600 python -m pytest tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=300 -x > gpurun_out/t_gen_pair.log 2>&1; echo "gen tests (pair) rc=$?"; tail -3 gpurun_out/t_gen_pair.log
for cfg in "1 0 2" "1 1 2" "0 0 2" "1 0 4" "1 0 1"; do
set -- $cfg
TG_FRAME_PAIR=$1 TG_FRAME_TAP=$2 timeout 300 python bench.py --steps 3 --warmup 3 --clips $3 --no-train --no-cpu-baseline --no-e2e > gpurun_out/bench_pair$1_tap$2_c$3.log 2>&1; echo "bench pair=$1 tap=$2 clips=$3 rc=$?"; tail -1 gpurun_out/bench_pair$1_tap$2_c$3.log | cut -c1-110; grep -o '"clocks": {[^}]*}' gpurun_out/bench_pair$1_tap$2_c$3.log
done
