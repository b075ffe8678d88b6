python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
tail -n 5 gpurun_out/t_all.log
python tools/probes/gpu_perf_probe.py 3840 > gpurun_out/perf.log 2>&1
head -32 gpurun_out/perf.log
