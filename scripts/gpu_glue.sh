#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_glue.py -m gpu -q -p no:cacheprovider --timeout=300 > gpurun_out/t_glue.log 2>&1; echo "glue tests rc=$?"; tail -2 gpurun_out/t_glue.log
python scripts/glue_bench.py > gpurun_out/glue_bench.json 2>gpurun_out/glue_bench.err; tail -3 gpurun_out/glue_bench.err
python -c "
import json
d=json.load(open('gpurun_out/glue_bench.json'))
print(d['workload'], d['peak'])
for k,v in d['kernels'].items(): print(k, round(v['us'],1), round(v['achieved']), round(v['frac'],3))
"
