# Evidence set of a build (GPU box, ONE GPU, ~7 minutes): pytest -m gpu, the three bench lines, the live kernel table, the ncu launch list
# of the bench command and the ncu metric sweeps.  Outputs land in gpurun_out/ (copy the summaries to profiles/).  usage: bash tools/final_round.sh <tag>
TAG=${1:-r02_v9}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; tail -n 2 gpurun_out/${TAG}_pytest.txt
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; head -c 300 gpurun_out/${TAG}_bench.json; echo
python bench.py --workload impala --no-cpu-baseline > gpurun_out/${TAG}_bench_impala.json 2>> gpurun_out/${TAG}_bench.err; head -c 200 gpurun_out/${TAG}_bench_impala.json; echo
python bench.py --network nature_cnn --no-cpu-baseline > gpurun_out/${TAG}_bench_nature_cnn.json 2>> gpurun_out/${TAG}_bench.err; head -c 200 gpurun_out/${TAG}_bench_nature_cnn.json; echo
python tools/probes/gpu_perf_probe.py 3840 > gpurun_out/${TAG}_kernel_table.txt 2>&1; head -n 3 gpurun_out/${TAG}_kernel_table.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 400 --csv --log-file gpurun_out/${TAG}_launch_list_raw.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launch_list_raw.csv "bench.py --steps 1 --warmup 3 --no-cpu-baseline" > gpurun_out/${TAG}_launch_list_summary.csv; head -n 8 gpurun_out/${TAG}_launch_list_summary.csv
bash tools/ncu_sweep.sh ${TAG} > /dev/null 2>&1; head -n 6 gpurun_out/${TAG}_ncu_sweep_learner_mb3840.txt
