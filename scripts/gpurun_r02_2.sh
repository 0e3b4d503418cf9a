cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_2.log
tail -5 gpurun_out/r02_pytest_2.log
python bench.py --only c1,c3,c4 --no-cpu-baseline > gpurun_out/r02_bench_v2_c134.json 2> gpurun_out/r02_bench_v2_c134.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_v2_c134.err
