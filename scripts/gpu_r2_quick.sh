#!/bin/bash
# GPU box: generator/conv parity tests + inference-only bench (value, roofline) — the fast loop for frame-kernel work.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python pytorch-tecogan_b200/build.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_generator.py tests/test_gpu_conv.py tests/test_gpu_pipeline.py ${EXTRA_TESTS} -m gpu -q -x -p no:cacheprovider --timeout=600 > gpurun_out/t_quick.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/t_quick.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-train --no-glue --no-cfg3 --no-cpu-baseline --no-torch-gpu > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'us/launch', round(d['roofline']['avg_launch_us'],1), d['clocks'])
PY
if [ -n "$TRACE" ]; then TG_FRAME_STAT_SEG=-1 TG_N=2 timeout 120 python scripts/frame_trace.py > gpurun_out/r02_trace_$TRACE.txt 2>&1; grep -E "^N=|res8|convT|ct2.0|ct3.2 128->128 c0|ct6|out 64" gpurun_out/r02_trace_$TRACE.txt | cut -c1-110; fi
