cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --only c2 --only-step > gpurun_out/r02_bench_n2_v6_g4.json 2> gpurun_out/r02_bench_n2_v6_g4.err; echo "g4 rc=$?"
tail -3 gpurun_out/r02_bench_n2_v6_g4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --only c2 --only-step --graph-steps 10 > gpurun_out/r02_bench_n2_v6_g10.json 2> gpurun_out/r02_bench_n2_v6_g10.err; echo "g10 rc=$?"
python bench.py --only c2 --only-step > gpurun_out/r02_bench_n1_v6_g4.json 2> gpurun_out/r02_bench_n1_v6_g4.err; echo "n1 g4 rc=$?"
python bench.py --only c2 --only-step --graph-steps 1 > gpurun_out/r02_bench_n1_v6_g1.json 2> gpurun_out/r02_bench_n1_v6_g1.err; echo "n1 g1 rc=$?"
