#!/bin/bash
# 2-GPU (gpurun --gpus 2): generator tests incl. the 4K case, DP parity test, bench at N=2 and N=1 on the same box
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=600 > gpurun_out/t_gen.log 2>&1; echo "gen rc=$?"; tail -2 gpurun_out/t_gen.log
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -p no:cacheprovider --timeout=800 -s > gpurun_out/t_dp.log 2>&1
echo "dp rc=$?"; tail -4 gpurun_out/t_dp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench2 rc=$?"; tail -3 gpurun_out/bench_n2.err; cut -c1-300 gpurun_out/bench_n2.json
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_samebox.json 2> gpurun_out/bench_n1.err
echo "bench1 rc=$?"; cut -c1-200 gpurun_out/bench_n1_samebox.json
