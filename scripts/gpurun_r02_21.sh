cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_effdet.py tests/test_golden.py tests/test_gpu_fullsize_properties.py tests/test_gpu_reference_emulated.py -m gpu -x -q 2>&1 | tail -2
python bench.py --only c4 --no-cpu-baseline > gpurun_out/r02_bench_v21_c4.json 2> gpurun_out/r02_bench_v21_c4.err; echo "bench rc=$?"
A="--only c4 --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
python bench.py $A > gpurun_out/plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_v21_c4.csv python bench.py $A > gpurun_out/ncu_c4.log 2>&1
