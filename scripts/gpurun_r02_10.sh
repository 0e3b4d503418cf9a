cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_10.log
tail -5 gpurun_out/r02_pytest_10.log
python bench.py --only c3 --no-cpu-baseline > gpurun_out/r02_bench_v10_c3.json 2> gpurun_out/r02_bench_v10_c3.err; echo "bench rc=$?"
A="--only c3 --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
python bench.py $A > gpurun_out/plain_c3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_v10_c3.csv python bench.py $A > gpurun_out/ncu_c3.log 2>&1
