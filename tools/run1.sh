python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
tail -n 5 gpurun_out/t_all.log
for f in 0 1; do echo "=== CLEANBA_FUSE_POOL_BWD=$f"; CLEANBA_FUSE_POOL_BWD=$f python tests/gpu_perf_probe.py 3840 2>&1 | grep -E "==|wgrad<cin4|pool_bwd@84|sum of"; done > gpurun_out/perf.log 2>&1
cat gpurun_out/perf.log
