#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=600 -x > gpurun_out/t_gen.log 2>&1; echo "gen tests rc=$?"; tail -3 gpurun_out/t_gen.log
for d in 0 4 5; do
TG_FRAME_DBG=$d timeout 600 python bench.py --steps 3 --warmup 3 --clips 2 --no-train --no-cpu-baseline --no-e2e > gpurun_out/bench_v6_dbg$d.log 2>&1; echo "bench v5 dbg=$d clips=2"; tail -1 gpurun_out/bench_v6_dbg$d.log | cut -c1-120; grep -o '"clocks": {[^}]*}' gpurun_out/bench_v6_dbg$d.log
done
for n in 2; do TG_N=$n timeout 300 python scripts/frame_trace.py > gpurun_out/frame_trace_v6_n$n.txt 2>&1; head -1 gpurun_out/frame_trace_v6_n$n.txt; done
TG_FRAME_DBG=4 TG_N=2 timeout 300 python scripts/frame_trace.py > gpurun_out/frame_trace_v6_nofence_n2.txt 2>&1; head -1 gpurun_out/frame_trace_v6_nofence_n2.txt
TG_FRAME_DBG=7 TG_N=2 timeout 300 python scripts/frame_trace.py > gpurun_out/frame_trace_v6_dbg7_n2.txt 2>&1; head -1 gpurun_out/frame_trace_v6_dbg7_n2.txt
