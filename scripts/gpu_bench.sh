#!/bin/bash
# bench run (gpurun): full default bench line + a quick train-only probe with kernel profile
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
