#!/bin/bash
# same-box A/B of warp-kernel builds (variant libraries built by the caller): arguments = variant names ("" = product)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for rep in 1 2; do
for v in "$@"; do
  lib=""; [ "$v" != "prod" ] && lib="pytorch-tecogan_b200/libtecogan_b200.$v.so"
  TG_GLUE_LIB=$lib python scripts/glue_bench.py > gpurun_out/glue_ab.json 2>gpurun_out/glue_ab.err || tail -n 3 gpurun_out/glue_ab.err
  python -c "
import json
d=json.load(open('gpurun_out/glue_ab.json'))
print('$v', ' '.join(f\"{k.split('_')[2]}={v['us']:.1f}us/{v['frac']:.3f}\" for k,v in d['kernels'].items() if 'warp_bilinear' in k))
"
done
done
