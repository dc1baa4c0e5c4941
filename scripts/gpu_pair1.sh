#!/bin/bash
# GPU box: pair-mode (cta_group::2) frame kernel - generator parity tests, then A/B bench against the single-CTA kernel.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=300 -x > gpurun_out/t_gen_pair.log 2>&1; echo "gen tests (pair) rc=$?"; tail -5 gpurun_out/t_gen_pair.log
for p in 1 0; do
TG_FRAME_PAIR=$p timeout 300 python bench.py --steps 3 --warmup 3 --clips 2 --no-train --no-cpu-baseline --no-e2e > gpurun_out/bench_pair$p.log 2>&1; echo "bench pair=$p rc=$?"; tail -1 gpurun_out/bench_pair$p.log | cut -c1-140; grep -o '"clocks": {[^}]*}' gpurun_out/bench_pair$p.log; grep -o '"roofline": {[^}]*}' gpurun_out/bench_pair$p.log | cut -c1-330
done
TG_N=2 timeout 300 python scripts/frame_trace.py > gpurun_out/frame_trace_pair_n2.txt 2>&1; head -1 gpurun_out/frame_trace_pair_n2.txt
