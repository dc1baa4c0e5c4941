#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=600 -x > gpurun_out/t_gen.log 2>&1
echo "gen tests rc=$?"; tail -30 gpurun_out/t_gen.log
timeout 300 python scripts/perf_probe.py > gpurun_out/perf_probe.log 2>&1; tail -4 gpurun_out/perf_probe.log
TG_N=2 timeout 300 python scripts/perf_probe.py > gpurun_out/perf_probe_n2.log 2>&1; tail -4 gpurun_out/perf_probe_n2.log
