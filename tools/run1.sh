set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
python tests/gpu_perf_probe.py 3840 > gpurun_out/perf.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_bytes.sum,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/sweep_learner.csv python tests/gpu_ncu_step.py 3840 learner > gpurun_out/sweep_learner.log 2>&1
timeout 300 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/sweep_actor.csv python tests/gpu_ncu_step.py 60 actor > gpurun_out/sweep_actor.log 2>&1
tail -3 gpurun_out/t_all.log; head -12 gpurun_out/perf.log
