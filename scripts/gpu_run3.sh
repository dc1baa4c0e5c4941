#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/mma_bench scripts/mma_bench.cu && timeout 120 /tmp/mma_bench > gpurun_out/mma_bench.log 2>&1
echo "mma_bench rc=$?"; cat gpurun_out/mma_bench.log
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=300 -k "not dx3" > gpurun_out/t_conv_gen.log 2>&1
echo "conv+gen rc=$?"; tail -8 gpurun_out/t_conv_gen.log
timeout 300 python scripts/perf_probe.py > gpurun_out/perf_probe_pdl.log 2>&1; cat gpurun_out/perf_probe_pdl.log
TG_PDL=0 timeout 300 python scripts/perf_probe.py > gpurun_out/perf_probe_nopdl.log 2>&1; cat gpurun_out/perf_probe_nopdl.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench.log
