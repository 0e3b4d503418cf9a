cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_9.log
tail -5 gpurun_out/r02_pytest_9.log
for v in 8 6 5 none; do
  if [ $v = none ]; then export B200_YL_NO_PIPE=1; else export B200_YL_PIPE_MINB=$v; fi
  python bench.py --only c2 --only-step --repeats 9 > gpurun_out/r02_bench_v9_c2_pipe$v.json 2> gpurun_out/r02_bench_v9_c2_pipe$v.err; echo "pipe $v rc=$?"
  unset B200_YL_NO_PIPE B200_YL_PIPE_MINB
done
A="--only c2 --only-step --no-graph --steps 3 --warmup 3 --repeats 1"
python bench.py $A > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_v9_c2.csv python bench.py $A > gpurun_out/ncu_c2.log 2>&1
python bench.py $A > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ignore_pipe -s 4 -c 1 -o gpurun_out/r02_prof_ignore_pipe_v1 -f python bench.py $A > gpurun_out/ncu_full_c2.log 2>&1
echo "full c2 rc=$?"
A="--only c3 --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
python bench.py $A > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:effdet_stream_kernel -s 3 -c 1 -o gpurun_out/r02_prof_stream_v3 -f python bench.py $A > gpurun_out/ncu_full_c3.log 2>&1
echo "full c3 rc=$?"
