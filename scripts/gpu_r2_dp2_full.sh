#!/bin/bash
# 2-GPU box: the NCCL data-parallel test and the driver's own N = 2 command (default bench flags)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 240 python -m pytest tests/test_gpu_dp.py -m gpu -q -s -p no:cacheprovider --timeout=200 > gpurun_out/t_dp.log 2>&1
echo "dp test rc=$?"; grep -E "rank|passed|failed|Error" gpurun_out/t_dp.log | tail -n 8
t0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$? wall=$(( $(date +%s) - t0 ))s"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n2.log').read().strip().splitlines()[-1])
    print('N=2 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))
    for k,v in d['train'].items(): print(k, 'clips/s', round(v['value'],1), 'ms/step', round(v['ms_per_step'],2), 'graph', v.get('cuda_graph'), 'e2e', round(v['e2e']['value'],1), 'finite', v['losses_finite'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_n2.err').read()[-1500:])
PY
