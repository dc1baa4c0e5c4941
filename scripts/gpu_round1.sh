#!/bin/bash
# GPU box: parity tests, MMA microbench, bench line, ncu launch list + one full capture of the conv kernel.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -x > gpurun_out/t_gpu.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/t_gpu.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/mma_bench scripts/mma_bench.cu && timeout 120 /tmp/mma_bench > gpurun_out/mma_bench.log 2>&1
echo "mma_bench rc=$?"; cat gpurun_out/mma_bench.log
timeout 300 python scripts/perf_probe.py > gpurun_out/perf_probe.log 2>&1; cat gpurun_out/perf_probe.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --clips 1 --frames 6 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -o gpurun_out/r01_conv_v1 python scripts/ncu_target.py > gpurun_out/ncu1.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu1.log
