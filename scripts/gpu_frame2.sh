#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=600 -x > gpurun_out/t_gen.log 2>&1
echo "gen tests rc=$?"; tail -3 gpurun_out/t_gen.log
for n in 1 2; do
  TG_N=$n timeout 200 python scripts/frame_trace.py > gpurun_out/trace_n$n.log 2>&1; grep -E "^N=|res7|convT|ct|out|conv.0" gpurun_out/trace_n$n.log | cut -c1-22,68-120
done
for dbg in 3; do
  echo "== TG_FRAME_DBG=$dbg"
  TG_FRAME_DBG=$dbg TG_N=2 timeout 200 python scripts/frame_trace.py > gpurun_out/trace_n2_dbg$dbg.log 2>&1; head -1 gpurun_out/trace_n2_dbg$dbg.log
  TG_FRAME_DBG=$dbg TG_N=1 timeout 200 python scripts/frame_trace.py > gpurun_out/trace_n1_dbg$dbg.log 2>&1; head -1 gpurun_out/trace_n1_dbg$dbg.log
done
