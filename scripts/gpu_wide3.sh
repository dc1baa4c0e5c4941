#!/bin/bash
# GPU box: frame kernel v4 (16 epilogue warps, batched item decode, prefetched dependency polls): parity + bench + traces.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=600 -x > gpurun_out/t_gen.log 2>&1
echo "gen tests rc=$?"; tail -4 gpurun_out/t_gen.log
for c in 1 2 4; do
  timeout 600 python bench.py --steps 3 --warmup 3 --clips $c --no-train --no-cpu-baseline --no-e2e > gpurun_out/bench_v4_c$c.log 2>&1; echo "bench v4 clips=$c rc=$?"; tail -1 gpurun_out/bench_v4_c$c.log | cut -c1-120;  grep -o '"clocks": {[^}]*}' gpurun_out/bench_v4_c$c.log
done
TG_FRAME_DBG=4 timeout 600 python bench.py --steps 3 --warmup 3 --clips 2 --no-train --no-cpu-baseline --no-e2e > gpurun_out/bench_v4_nofence_c2.log 2>&1; echo "bench v4 nofence clips=2"; tail -1 gpurun_out/bench_v4_nofence_c2.log | cut -c1-120
TG_FRAME_WIDE=0 timeout 600 python bench.py --steps 3 --warmup 3 --clips 2 --no-train --no-cpu-baseline --no-e2e > gpurun_out/bench_v4_tall_c2.log 2>&1; echo "bench v4 tall clips=2"; tail -1 gpurun_out/bench_v4_tall_c2.log | cut -c1-120
for n in 1 2 4; do TG_N=$n timeout 300 python scripts/frame_trace.py > gpurun_out/frame_trace_v4_n$n.txt 2>&1; head -1 gpurun_out/frame_trace_v4_n$n.txt; done
