#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
for cfg in "1 0" "0 0" "1 64" "0 64"; do
  set -- $cfg
  TG_FRAME_PAIR=$1 TG_FRAME_DBG=$2 TG_N=2 timeout 120 python scripts/frame_trace.py > gpurun_out/trace_pair$1_dbg$2.txt 2>&1
  echo "== pair=$1 dbg=$2"; grep -E "total|^stat" gpurun_out/trace_pair$1_dbg$2.txt
done
