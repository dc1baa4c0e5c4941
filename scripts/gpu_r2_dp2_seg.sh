#!/bin/bash
# 2-GPU box: data-parallel test with the segmented graph capture, then the cfg5 train leg with it on / off
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 200 python -m pytest tests/test_gpu_dp.py -m gpu -q -s -p no:cacheprovider --timeout=180 > gpurun_out/t_dp.log 2>&1
echo "dp test rc=$?"; grep -E "rank|passed|failed|Error|error" gpurun_out/t_dp.log | tail -n 10
for v in 1 0; do
TG_TRAIN_GRAPH_DP=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$v bench.py --gpus 2 --steps 2 --warmup 3 --frames 10 --no-glue --no-cfg3 --no-e2e --no-cpu-baseline --no-torch-gpu > gpurun_out/bench_n2_dp$v.log 2> gpurun_out/bench_n2_dp$v.err
echo "TG_TRAIN_GRAPH_DP=$v bench rc=$?"
V=$v python - <<'PY'
import json, os
v=os.environ['V']
try:
    d=json.loads(open(f'gpurun_out/bench_n2_dp{v}.log').read().strip().splitlines()[-1])
    for k,x in d['train'].items(): print(k, 'clips/s', round(x['value'],1), 'ms/step', round(x['ms_per_step'],2), 'graph', x.get('cuda_graph'), 'e2e', round(x['e2e']['value'],1), 'finite', x['losses_finite'], x.get('allreduce',{}).get('exposed_ms_per_step'))
except Exception as e:
    print('parse failed', e); print(open(f'gpurun_out/bench_n2_dp{v}.err').read()[-2500:])
PY
done
