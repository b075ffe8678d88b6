# Round profile (GPU box, one GPU): bench line, ncu launch list of the bench command, ncu metric sweep of one learner minibatch
# and one actor step.  Outputs land in gpurun_out/ (copy the summaries to profiles/).
TAG=${1:-r01_v9}
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 400 --csv --log-file gpurun_out/${TAG}_launch_list_raw.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_bytes.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_sweep_learner_raw.csv \
    python tools/probes/gpu_ncu_step.py 3840 learner > /dev/null 2>&1
python tools/sweep_table.py gpurun_out/${TAG}_sweep_learner_raw.csv > gpurun_out/${TAG}_ncu_sweep_learner_mb3840.txt
timeout 300 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_sweep_actor_raw.csv \
    python tools/probes/gpu_ncu_step.py 60 actor > /dev/null 2>&1
python tools/sweep_table.py gpurun_out/${TAG}_sweep_actor_raw.csv > gpurun_out/${TAG}_ncu_sweep_actor_n60.txt
head -12 gpurun_out/${TAG}_ncu_sweep_learner_mb3840.txt
