cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_22.log
tail -4 gpurun_out/r02_pytest_22.log
python bench.py --only c1 --no-cpu-baseline > gpurun_out/r02_bench_v22_c1.json 2> gpurun_out/r02_bench_v22_c1.err; echo "bench rc=$?"
python bench.py --only c5 --c5-global-batch 64 --no-cpu-baseline > gpurun_out/r02_bench_v22_c5_b64.json 2> gpurun_out/r02_bench_v22_c5_b64.err; echo "bench rc=$?"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
A="--only c1 --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
python bench.py $A > gpurun_out/plain_c1.log 2>&1 && \
ncu --metrics $M --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_v22_c1.csv python bench.py $A > gpurun_out/ncu_c1.log 2>&1
python scripts/run_one.py c1_b1 6 > gpurun_out/plain_c1b1.log 2>&1 && \
ncu --metrics $M --clock-control none -c 100 --csv --log-file gpurun_out/r02_launches_v22_c1_b1.csv python scripts/run_one.py c1_b1 6 > gpurun_out/ncu_c1b1.log 2>&1
