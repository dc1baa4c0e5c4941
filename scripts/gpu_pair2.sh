#!/bin/bash
# GPU box: where does the pair-mode frame kernel lose time?  Segment traces under the measurement knobs.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
sel='total|res8.0|ct2.0|ct3.2 128->128 c0|convT128 c0|ct6|out 64'
for cfg in "1 0" "0 0" "1 8" "1 16" "1 24" "1 32" "1 1" "1 3"; do
  set -- $cfg
  TG_FRAME_PAIR=$1 TG_FRAME_DBG=$2 TG_N=2 timeout 120 python scripts/frame_trace.py > gpurun_out/trace_pair$1_dbg$2.txt 2>&1
  echo "== pair=$1 dbg=$2"; grep -E "$sel" gpurun_out/trace_pair$1_dbg$2.txt | cut -c1-22,60-90
done
