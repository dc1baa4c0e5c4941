#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -x > gpurun_out/t_gpu.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/t_gpu.log
for pr in 1; do
  echo "== TG_CONV_PAIR=$pr"
  TG_CONV_PAIR=$pr TG_STEPS=10 python scripts/train_probe.py 2>&1 | tail -1
  TG_CONV_PAIR=$pr TG_CFG=5 TG_STEPS=4 python scripts/train_probe.py 2>&1 | tail -1
done
