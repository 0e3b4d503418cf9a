cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_effdet.py tests/test_golden.py tests/test_gpu_yolo_decode.py tests/test_gpu_core.py tests/test_gpu_fullsize_properties.py -m gpu -x -q > gpurun_out/r02_pytest_5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_5.log
tail -5 gpurun_out/r02_pytest_5.log
python bench.py --only c1,c3,c4 --no-cpu-baseline > gpurun_out/r02_bench_v5_c134.json 2> gpurun_out/r02_bench_v5_c134.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_v5_c134.err
for c in c3 c4; do
  A="--only $c --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
  python bench.py $A > gpurun_out/plain_$c.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_v5_$c.csv python bench.py $A > gpurun_out/ncu_$c.log 2>&1
  echo "$c rc=$?"
done
A="--only c2 --only-step --no-graph --steps 3 --warmup 3 --repeats 1"
python bench.py $A > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:yolo_loss_ignore -s 4 -c 1 -o gpurun_out/r02_prof_ignore_v1 -f python bench.py $A > gpurun_out/ncu_full_c2.log 2>&1
echo "full c2 rc=$?"
python scripts/run_one.py c1_b1 6 > gpurun_out/plain_c1b1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:yolo_nms_finalize -s 4 -c 1 -o gpurun_out/r02_prof_nms_b1_v1 -f python scripts/run_one.py c1_b1 6 > gpurun_out/ncu_full_c1b1.log 2>&1
echo "full c1b1 rc=$?"
