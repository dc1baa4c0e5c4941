#!/bin/bash
# GPU box: backward / train parity tests + train-only bench legs (cfg4, cfg5).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
timeout 1200 python -m pytest tests/test_gpu_backward.py tests/test_gpu_train.py tests/test_gpu_optim.py tests/test_gpu_discriminator.py tests/test_gpu_generator.py -m gpu -q -p no:cacheprovider --timeout=900 -k "${TESTK:-not nothing}" > gpurun_out/t_train.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed" gpurun_out/t_train.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/t_train.log | head
timeout 900 python bench.py --steps 2 --warmup 3 --frames 4 --no-glue --no-cfg3 --no-cpu-baseline --no-torch-gpu --no-e2e > gpurun_out/bench_train.log 2> gpurun_out/bench_train.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_train.log').read().strip().splitlines()[-1])
for k,v in d['train'].items(): print(k, 'clips/s', round(v['value'],1), 'ms/step', round(v['ms_per_step'],2), 'TF/s', round(v['step_tflops'],1), 'graph', v.get('cuda_graph'), 'e2e', round(v['e2e']['value'],1), 'finite', v['losses_finite'])
PY
