cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_8.log
tail -5 gpurun_out/r02_pytest_8.log
python bench.py --only c1,c3,c4 --no-cpu-baseline > gpurun_out/r02_bench_v8_c134.json 2> gpurun_out/r02_bench_v8_c134.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_v8_c134.err
python bench.py --only c5 --c5-global-batch 64 --no-cpu-baseline > gpurun_out/r02_bench_v8_c5_b64.json 2> gpurun_out/r02_bench_v8_c5_b64.err; echo "bench rc=$?"
for c in c3 c4; do
  A="--only $c --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
  python bench.py $A > gpurun_out/plain_$c.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_v8_$c.csv python bench.py $A > gpurun_out/ncu_$c.log 2>&1
  echo "$c rc=$?"
done
A="--only c3 --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
python bench.py $A > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:effdet_stream -s 3 -c 1 -o gpurun_out/r02_prof_stream_v3 -f python bench.py $A > gpurun_out/ncu_full_c3.log 2>&1
echo "full c3 rc=$?"
