#!/bin/bash
# D backward parity + the backward / discriminator suites (gpurun)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_discriminator.py -m gpu -q -p no:cacheprovider --timeout=600 -x > gpurun_out/t_disc.log 2>&1
echo "disc rc=$?"; tail -25 gpurun_out/t_disc.log
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=600 > gpurun_out/t_bwd_gen.log 2>&1
echo "bwd+gen rc=$?"; tail -8 gpurun_out/t_bwd_gen.log
