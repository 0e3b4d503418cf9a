cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_6.log
tail -5 gpurun_out/r02_pytest_6.log
python bench.py --only c1,c4 --no-cpu-baseline > gpurun_out/r02_bench_v6_c14.json 2> gpurun_out/r02_bench_v6_c14.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_v6_c14.err
python bench.py --only c5 --c5-global-batch 64 --no-cpu-baseline > gpurun_out/r02_bench_v6_c5_b64.json 2> gpurun_out/r02_bench_v6_c5_b64.err; echo "bench rc=$?"
for c in c1 c4; do
  A="--only $c --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
  python bench.py $A > gpurun_out/plain_$c.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_v6_$c.csv python bench.py $A > gpurun_out/ncu_$c.log 2>&1
  echo "$c rc=$?"
done
python scripts/run_one.py c1_b1 6 > gpurun_out/plain_c1b1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r02_launches_v6_c1b1.csv python scripts/run_one.py c1_b1 6 > gpurun_out/ncu_c1b1.log 2>&1
python scripts/run_one.py c1_b1 6 > gpurun_out/plain_c1b1.log 2>&1 && \
ncu --set full --sampling-interval 0 --clock-control none --import-source on -k regex:yolo_nms_finalize -s 4 -c 1 -o gpurun_out/r02_prof_nms_b1_v2 -f python scripts/run_one.py c1_b1 6 > gpurun_out/ncu_full_c1b1.log 2>&1
echo "full c1b1 rc=$?"
