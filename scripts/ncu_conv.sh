mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -o gpurun_out/r01_conv_v1 python scripts/ncu_target.py > gpurun_out/ncu1.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu1.log
TG_N=2 timeout 200 python scripts/perf_probe.py > gpurun_out/perf_probe_n2.log 2>&1; cat gpurun_out/perf_probe_n2.log
TG_N=4 timeout 200 python scripts/perf_probe.py > gpurun_out/perf_probe_n4.log 2>&1; cat gpurun_out/perf_probe_n4.log
