#!/bin/bash
# GPU box: whole-frame timeline of the frame kernel under measurement knobs.  usage: gpu_r2_dbg.sh "DBG ..."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
for dbg in ${1:-0 16}; do
  TG_FRAME_DBG=$dbg TG_FRAME_STAT_SEG=-1 TG_N=2 timeout 120 python scripts/frame_trace.py > gpurun_out/r02_all_dbg$dbg.txt 2>&1
  echo "== dbg=$dbg"; grep -E "^N=|res8|convT|ct2.0|ct3.2 128->128 c0|ct6|out 64|^stat" gpurun_out/r02_all_dbg$dbg.txt | cut -c1-110
done
