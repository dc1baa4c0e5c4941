#!/bin/bash
# GPU box: glue parity tests + HBM roofline of the glue kernels at 4K, branch-free warp kernel on / off (same box).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1 || { tail -n 20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_glue.py -m gpu -q -p no:cacheprovider --timeout=300 > gpurun_out/t_glue.log 2>&1; echo "glue tests rc=$?"; tail -n 3 gpurun_out/t_glue.log
for v in 1 0 1 0; do
TG_WARP_BRANCHFREE=$v python scripts/glue_bench.py > gpurun_out/glue_bench_bf$v.json 2>gpurun_out/glue_bench.err; tail -n 3 gpurun_out/glue_bench.err
python -c "
import json
d=json.load(open('gpurun_out/glue_bench_bf$v.json'))
print('branch_free=$v', d['workload'], d['peak'])
for k,v in d['kernels'].items():
    if 'warp' in k or 'copy' in k: print(' ', k, round(v['us'],1), round(v['achieved']), round(v['frac'],3))
"
done
