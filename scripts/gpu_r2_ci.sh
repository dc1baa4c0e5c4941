#!/bin/bash
# GPU box (round 2): full parity suite (no -x, slowest tests listed), smoke, default bench + reference arm.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt 2>&1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 --durations=20 -s > gpurun_out/t_gpu.log 2>&1
echo "gpu tests rc=$?"; grep -E "passed|failed|error" gpurun_out/t_gpu.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/t_gpu.log | head -40
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
if [ "$1" != "nobench" ]; then
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2.log 2> gpurun_out/bench_r2.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_r2.log | cut -c1-400
fi
