#!/bin/bash
# GPU box: full parity suite (no -x), smoke, default bench + reference arm, ncu launch list + full capture of the pair-mode frame kernel.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 > gpurun_out/t_gpu.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/t_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v10.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_v10.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_frame_v10.csv python bench.py --steps 1 --warmup 3 --clips 2 --frames 8 --no-e2e --no-cpu-baseline --no-train --no-glue > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 4 -c 1 -f -o gpurun_out/r01_frame_v10 python bench.py --steps 1 --warmup 3 --clips 2 --frames 2 --no-e2e --no-cpu-baseline --no-train --no-glue > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
