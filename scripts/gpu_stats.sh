#!/bin/bash
# GPU box: stall accounting of the frame kernel.  usage: gpu_stats.sh "PAIR TAP SEG" ...
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
for cfg in "$@"; do
  set -- $cfg
  TG_FRAME_STAT_SEG=$3 TG_FRAME_PAIR=$1 TG_FRAME_TAP=$2 TG_N=2 timeout 120 python scripts/frame_trace.py > gpurun_out/stats_$1_$2_$3.txt 2>&1
  echo "== pair=$1 tap=$2 seg=$3"; grep -E "total|res8.0|ct2.0|^stat" gpurun_out/stats_$1_$2_$3.txt | cut -c1-100
done
