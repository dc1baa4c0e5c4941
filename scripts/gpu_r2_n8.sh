#!/bin/bash
# N-GPU box: the driver's own scaling command (default bench flags) at N = $1 (default 8), then the reference arm the same way.
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; exit 1; }
t0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "bench n$N rc=$? wall=$(( $(date +%s) - t0 ))s"
t0=$(date +%s)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2> gpurun_out/bench_ref_n$N.err
echo "ref n$N rc=$? wall=$(( $(date +%s) - t0 ))s lines=$(grep -c impl gpurun_out/bench_ref_n$N.log)"
N=$N python - <<'PY'
import json, os
N=os.environ['N']
try:
    d=json.loads(open(f'gpurun_out/bench_n{N}.log').read().strip().splitlines()[-1])
    print('N',d['n_gpus'],'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'clocks', d['clocks'])
    for k,v in d['train'].items(): print(k, 'clips/s', round(v['value'],1), 'ms/step', round(v['ms_per_step'],2), 'graph', v.get('cuda_graph'), 'e2e', round(v['e2e']['value'],1), 'finite', v['losses_finite'], v.get('allreduce'))
except Exception as e:
    print('parse failed', e); print(open(f'gpurun_out/bench_n{N}.err').read()[-2500:])
PY
