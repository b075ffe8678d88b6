# ncu metric sweep of ONE learner minibatch step (mb 3840) and one actor step (n 60); one GPU, ~2 minutes.  TAG = output prefix.
TAG=${1:-r02}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_bytes.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_sweep_learner_raw.csv \
    python tools/probes/gpu_ncu_step.py 3840 learner > /dev/null 2>&1
python tools/sweep_table.py gpurun_out/${TAG}_sweep_learner_raw.csv > gpurun_out/${TAG}_ncu_sweep_learner_mb3840.txt
timeout 300 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_sweep_actor_raw.csv \
    python tools/probes/gpu_ncu_step.py 60 actor > /dev/null 2>&1
python tools/sweep_table.py gpurun_out/${TAG}_sweep_actor_raw.csv > gpurun_out/${TAG}_ncu_sweep_actor_n60.txt
cat gpurun_out/${TAG}_ncu_sweep_learner_mb3840.txt gpurun_out/${TAG}_ncu_sweep_actor_n60.txt
