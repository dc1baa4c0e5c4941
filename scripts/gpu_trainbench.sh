#!/bin/bash
# GPU box: training tests, then the train legs of bench.py with the fused / plain stock Adam
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_discriminator.py tests/test_gpu_backward.py -m gpu -q -p no:cacheprovider --timeout=600 > gpurun_out/t_train.log 2>&1; echo "train tests rc=$?"; tail -2 gpurun_out/t_train.log
for f in 1 0; do
TG_BENCH_FUSED_ADAM=$f timeout 600 python bench.py --steps 2 --warmup 3 --no-glue --no-cpu-baseline --no-e2e > gpurun_out/bench_train_f$f.log 2>&1; echo "fused=$f rc=$?"
tail -1 gpurun_out/bench_train_f$f.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d['train'].items(): print(k, round(v['value'],1), round(v['ms_per_step'],2), round(v['e2e']['value'],1), v['gpu_launches'])
"
done
