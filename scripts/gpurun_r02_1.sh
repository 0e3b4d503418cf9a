set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_1.log
tail -5 gpurun_out/r02_pytest_1.log
python bench.py --verbose > gpurun_out/r02_bench_n1_v1.json 2> gpurun_out/r02_bench_n1_v1.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_n1_v1.err
for c in c2 c5 c1 c3 c4; do
  A="--only $c --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
  python bench.py $A > gpurun_out/plain_$c.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_v1_$c.csv python bench.py $A > gpurun_out/ncu_$c.log 2>&1
  echo "$c rc=$?"
done
