#!/bin/bash
# GPU box: what bounds a segment of the frame kernel?  usage: gpu_r2_outconv.sh SEG "DBG ..."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
SEG=${1:-43}
for dbg in ${2:-0 512 1024 1536}; do
  TG_FRAME_DBG=$dbg TG_FRAME_STAT_SEG=$SEG TG_N=2 timeout 120 python scripts/frame_trace.py > gpurun_out/r02_seg${SEG}_dbg$dbg.txt 2>&1
  echo "== seg=$SEG dbg=$dbg"; grep -E "^N=|res8|ct2.0|ct6|out 64|^stat" gpurun_out/r02_seg${SEG}_dbg$dbg.txt | cut -c1-110
done
