cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_yolo_loss.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
for v in "0 2048" "10 512" "10 2048" "10 8192" "12 2048"; do
  set -- $v
  B200_YL_SPLIT=$1 B200_YL_EXACT_CTAS=$2 python bench.py --only c2 --only-step --repeats 9 > gpurun_out/r02_bench_v17_c2_split$1_$2.json 2> gpurun_out/r02_bench_v17_c2_split$1_$2.err; echo "split $v rc=$?"
done
A="--only c2 --only-step --no-graph --steps 3 --warmup 3 --repeats 1"
B200_YL_SPLIT=10 python bench.py $A > gpurun_out/plain_c2.log 2>&1 && \
B200_YL_SPLIT=10 ncu --set full --clock-control none --import-source on -k regex:ignore_exact -s 4 -c 1 -o gpurun_out/r02_prof_ignore_exact_v1 -f python bench.py $A > gpurun_out/ncu_full_c2.log 2>&1
