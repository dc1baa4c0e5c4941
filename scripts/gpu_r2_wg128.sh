#!/bin/bash
# GPU box: backward / train / discriminator parity tests, then cfg4 / cfg5 step time with the 128-channel ky-stacked wgrad on / off.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -n 30 gpurun_out/build.log; exit 1; }
timeout 1200 python -m pytest tests/test_gpu_backward.py tests/test_gpu_train.py tests/test_gpu_discriminator.py tests/test_gpu_perceptual.py -m gpu -x -q -p no:cacheprovider --timeout=600 > gpurun_out/t_bwd.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed" gpurun_out/t_bwd.log | tail -n 2; grep -E "^(FAILED|ERROR)|Error|error" gpurun_out/t_bwd.log | head -n 12
for rep in 1 2; do
for v in 1 0; do
  echo "TG_WGRAD_KY128=$v: $(TG_WGRAD_KY128=$v TG_STEPS=10 TG_WARM=4 timeout 300 python scripts/train_probe.py 2>&1 | tail -n 1)"
  echo "TG_WGRAD_KY128=$v: $(TG_WGRAD_KY128=$v TG_CFG=5 TG_STEPS=5 TG_WARM=4 timeout 300 python scripts/train_probe.py 2>&1 | tail -n 1)"
done
done
