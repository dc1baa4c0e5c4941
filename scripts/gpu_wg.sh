#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
for t in 2 4; do
  echo "== tiles per slab $t"
  TG_WGRAD_TILES_PER_SLAB=$t TG_STEPS=10 python scripts/train_probe.py 2>&1 | tail -1
  TG_WGRAD_TILES_PER_SLAB=$t TG_CFG=5 TG_STEPS=4 python scripts/train_probe.py 2>&1 | tail -1
done
echo "== single-CTA frame kernel"
TG_FRAME_PAIR=0 TG_STAGES=1 python scripts/train_probe.py 2>&1 | tail -9
TG_FRAME_PAIR=0 TG_CFG=5 TG_STEPS=4 python scripts/train_probe.py 2>&1 | tail -1
