#!/bin/bash
# training-step parity (gpurun)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider --timeout=600 -s > gpurun_out/t_train.log 2>&1
echo "train rc=$?"; tail -40 gpurun_out/t_train.log
timeout 900 python -m pytest tests/test_gpu_discriminator.py tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=600 -s > gpurun_out/t_disc_gen.log 2>&1
echo "disc+gen rc=$?"; tail -12 gpurun_out/t_disc_gen.log
