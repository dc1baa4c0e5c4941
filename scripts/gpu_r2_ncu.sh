#!/bin/bash
# GPU box, round 2 evidence pass: ncu launch lists (inference bench command, eager cfg4 / cfg5 training steps) and
# `--set full` captures of the frame kernel and of the ky-stacked wgrad kernel.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
NOLEGS="--no-e2e --no-cpu-baseline --no-train --no-glue --no-cfg3 --no-torch-gpu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/r02_ncu_launches_frame.csv \
  python bench.py --steps 1 --warmup 3 --clips 2 --frames 8 $NOLEGS > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 4 -c 1 -f -o gpurun_out/r02_frame_final \
  python bench.py --steps 1 --warmup 3 --clips 2 --frames 2 $NOLEGS > gpurun_out/ncu_full.log 2>&1
echo "ncu full frame rc=$?"; tail -2 gpurun_out/ncu_full.log
TG_TRAIN_GRAPH=0 TG_CFG=5 TG_STEPS=1 TG_WARM=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r02_ncu_launches_train_cfg5.csv python scripts/train_probe.py > gpurun_out/ncu_train_cfg5.log 2>&1
echo "ncu cfg5 rc=$?"
TG_TRAIN_GRAPH=0 TG_STEPS=1 TG_WARM=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r02_ncu_launches_train_cfg4.csv python scripts/train_probe.py > gpurun_out/ncu_train_cfg4.log 2>&1
echo "ncu cfg4 rc=$?"
TG_TRAIN_GRAPH=0 TG_CFG=5 TG_STEPS=1 TG_WARM=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad3x3_ky \
  -s 40 -c 2 -f -o gpurun_out/r02_wgrad_ky_cfg5 python scripts/train_probe.py > gpurun_out/ncu_wgrad.log 2>&1
echo "ncu full wgrad rc=$?"; tail -2 gpurun_out/ncu_wgrad.log
TG_TRAIN_GRAPH=0 TG_CFG=5 TG_STEPS=1 TG_WARM=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgrad\|conv_tc \
  -s 200 -c 2 -f -o gpurun_out/r02_conv_cfg5 python scripts/train_probe.py > gpurun_out/ncu_conv.log 2>&1
echo "ncu full conv rc=$?"; tail -2 gpurun_out/ncu_conv.log
TG_STAGES=1 TG_CFG=5 TG_TRAIN_GRAPH=0 timeout 300 python scripts/train_probe.py 2>&1 | tail -12
