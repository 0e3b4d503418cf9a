cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_effdet.py tests/test_golden.py tests/test_gpu_yolo_decode.py -m gpu -x -q > gpurun_out/r02_pytest_4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_4.log
tail -5 gpurun_out/r02_pytest_4.log
python bench.py --only c3,c4,c5 --no-cpu-baseline > gpurun_out/r02_bench_v4_c345.json 2> gpurun_out/r02_bench_v4_c345.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_v4_c345.err
python bench.py --only c5 --c5-global-batch 64 --no-cpu-baseline > gpurun_out/r02_bench_v4_c5_b64.json 2> gpurun_out/r02_bench_v4_c5_b64.err; echo "bench rc=$?"
for c in c3 c4; do
  A="--only $c --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
  python bench.py $A > gpurun_out/plain_$c.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_v4_$c.csv python bench.py $A > gpurun_out/ncu_$c.log 2>&1
  echo "$c rc=$?"
done
