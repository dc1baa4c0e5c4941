#!/bin/bash
# GPU box: quick A/B of frame-kernel variants: generator parity tests, then value-only bench lines.  usage: gpu_ab.sh "PAIR TAP CLIPS" ...
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=300 -x > gpurun_out/t_gen.log 2>&1; echo "gen tests rc=$?"; tail -2 gpurun_out/t_gen.log
for cfg in "$@"; do
set -- $cfg
TG_FRAME_PAIR=$1 TG_FRAME_TAP=$2 timeout 300 python bench.py --steps 5 --warmup 3 --clips $3 --no-train --no-cpu-baseline --no-e2e > gpurun_out/bench_ab_$1_$2_$3.log 2>&1; echo "bench pair=$1 tap=$2 clips=$3 rc=$?"; tail -1 gpurun_out/bench_ab_$1_$2_$3.log | cut -c50-110; grep -o '"clocks": {[^}]*}' gpurun_out/bench_ab_$1_$2_$3.log; grep -o '"avg_launch_us": [0-9.]*' gpurun_out/bench_ab_$1_$2_$3.log
done
