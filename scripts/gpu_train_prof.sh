#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
TG_CFG=5 TG_STEPS=1 TG_WARM=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_train_cfg5.csv python scripts/train_probe.py > gpurun_out/ncu_train_cfg5.log 2>&1
echo "ncu cfg5 rc=$?"
TG_STEPS=1 TG_WARM=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_train_cfg4.csv python scripts/train_probe.py > gpurun_out/ncu_train_cfg4.log 2>&1
echo "ncu cfg4 rc=$?"
TG_STAGES=1 TG_CFG=5 python scripts/train_probe.py 2>&1 | tail -9
