#!/bin/bash
# GPU box: full parity suite, bench at 1/2/4 clips per step, ncu launch list + full capture of the frame kernel.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -x > gpurun_out/t_gpu.log 2>&1
echo "gpu tests rc=$?"; tail -3 gpurun_out/t_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
for c in 1 2 4; do
  timeout 600 python bench.py --steps 3 --warmup 3 --clips $c > gpurun_out/bench_c$c.log 2>&1; echo "bench clips=$c rc=$?"; tail -1 gpurun_out/bench_c$c.log | cut -c1-260
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_frame.csv python bench.py --steps 1 --warmup 3 --clips 2 --frames 8 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 4 -c 1 -o gpurun_out/r01_frame_v2 python bench.py --steps 1 --warmup 3 --clips 2 --frames 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
