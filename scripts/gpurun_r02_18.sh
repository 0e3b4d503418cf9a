cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_yolo_loss.py tests/test_golden.py tests/test_gpu_reference_emulated.py tests/test_gpu_fullsize_properties.py -m gpu -x -q 2>&1 | tail -2
for v in "0 2048" "10 1024" "10 2048" "12 2048" "8 2048"; do
  set -- $v
  B200_YL_SPLIT=$1 B200_YL_EXACT_CTAS=$2 python bench.py --only c2 --only-step --repeats 9 > gpurun_out/r02_bench_v18_c2_split$1_$2.json 2> gpurun_out/r02_bench_v18_c2_split$1_$2.err; echo "split $v rc=$?"
done
A="--only c2 --only-step --no-graph --steps 3 --warmup 3 --repeats 1"
B200_YL_SPLIT=10 python bench.py $A > gpurun_out/plain_c2.log 2>&1 && \
B200_YL_SPLIT=10 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_v18_c2.csv python bench.py $A > gpurun_out/ncu_c2.log 2>&1
