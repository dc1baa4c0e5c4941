#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests file by file in separate processes (a trapped
# kernel poisons the CUDA context of its process only), then a perf probe.  Logs -> gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_glue.py -m gpu -q -p no:cacheprovider --timeout=300 > gpurun_out/t_glue.log 2>&1
echo "glue rc=$?"; tail -5 gpurun_out/t_glue.log
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider --timeout=300 -k "dx3" > gpurun_out/t_conv_dx3.log 2>&1
echo "conv dx3 rc=$?"; tail -5 gpurun_out/t_conv_dx3.log
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider --timeout=300 -k "not dx3" > gpurun_out/t_conv_halo.log 2>&1
echo "conv halo rc=$?"; tail -5 gpurun_out/t_conv_halo.log
timeout 900 python -m pytest tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=600 > gpurun_out/t_gen.log 2>&1
echo "gen rc=$?"; tail -5 gpurun_out/t_gen.log
timeout 300 python scripts/perf_probe.py > gpurun_out/perf_probe.log 2>&1
echo "probe rc=$?"; cat gpurun_out/perf_probe.log
