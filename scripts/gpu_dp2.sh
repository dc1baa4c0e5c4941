#!/bin/bash
# 2-GPU: DP parity test + bench at N=2 (gpurun --gpus 2)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -p no:cacheprovider --timeout=800 -s > gpurun_out/t_dp.log 2>&1
echo "dp rc=$?"; tail -15 gpurun_out/t_dp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench2 rc=$?"; tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
