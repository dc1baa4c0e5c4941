#!/bin/bash
# GPU box: wide-geometry frame kernel (3 filter columns fused into one N=192 MMA): parity, A/B bench, trace.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=600 -x > gpurun_out/t_gen.log 2>&1
echo "gen tests rc=$?"; tail -15 gpurun_out/t_gen.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
for c in 1 2; do
  timeout 600 python bench.py --steps 3 --warmup 3 --clips $c --no-train --no-cpu-baseline > gpurun_out/bench_wide_c$c.log 2>&1; echo "bench wide clips=$c rc=$?"; tail -1 gpurun_out/bench_wide_c$c.log | cut -c1-200
  TG_FRAME_WIDE=0 timeout 600 python bench.py --steps 3 --warmup 3 --clips $c --no-train --no-cpu-baseline --no-e2e > gpurun_out/bench_tall_c$c.log 2>&1; echo "bench tall clips=$c rc=$?"; tail -1 gpurun_out/bench_tall_c$c.log | cut -c1-200
done
timeout 300 python scripts/frame_trace.py > gpurun_out/frame_trace_wide_n1.txt 2>&1; echo "trace rc=$?"; head -1 gpurun_out/frame_trace_wide_n1.txt
TG_FRAME_WIDE=0 timeout 300 python scripts/frame_trace.py > gpurun_out/frame_trace_tall_n1.txt 2>&1; head -1 gpurun_out/frame_trace_tall_n1.txt
