cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_16.log
tail -5 gpurun_out/r02_pytest_16.log
for v in 0 8 10 12 16; do
  B200_YL_SPLIT=$v python bench.py --only c2 --only-step --repeats 9 > gpurun_out/r02_bench_v16_c2_split$v.json 2> gpurun_out/r02_bench_v16_c2_split$v.err; echo "split $v rc=$?"
done
A="--only c2 --only-step --no-graph --steps 3 --warmup 3 --repeats 1"
python bench.py $A > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_v16_c2.csv python bench.py $A > gpurun_out/ncu_c2.log 2>&1
