# Round-2 final single-GPU run: launch lists of every config (-> profiles/r02_kernels.json), then tests, smoke, both bench arms.
cd $GRAFT_REPO_ROOT
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for c in c2 c5 c1 c3 c4; do
  A="--only $c --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
  python bench.py $A > gpurun_out/plain_$c.log 2>&1 && \
  ncu --metrics $M --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_final_$c.csv python bench.py $A > gpurun_out/ncu_$c.log 2>&1
  echo "$c rc=$?"
done
A="--only c5 --c5-global-batch 64 --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
python bench.py $A > gpurun_out/plain_c5b64.log 2>&1 && \
ncu --metrics $M --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_final_c5_b64.csv python bench.py $A > gpurun_out/ncu_c5b64.log 2>&1
python scripts/run_one.py c1_b1 6 > gpurun_out/plain_c1b1.log 2>&1 && \
ncu --metrics $M --clock-control none -c 100 --csv --log-file gpurun_out/r02_launches_final_c1_b1.csv python scripts/run_one.py c1_b1 6 > gpurun_out/ncu_c1b1.log 2>&1
G=gpurun_out
python profiles/summarize_launches.py c2=$G/r02_launches_final_c2.csv:7:64 c5=$G/r02_launches_final_c5.csv:7:512 c1=$G/r02_launches_final_c1.csv:6:256 c3=$G/r02_launches_final_c3.csv:6:128 c4=$G/r02_launches_final_c4.csv:6:16 c5_b64=$G/r02_launches_final_c5_b64.csv:7:64 c1_b1=$G/r02_launches_final_c1_b1.csv:6:1 -o profiles/r02_kernels.json > gpurun_out/r02_kernels_summary.txt
cp profiles/r02_kernels.json gpurun_out/r02_kernels.json
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_final.log
tail -4 gpurun_out/r02_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$?"
python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
python bench.py --only c5 --c5-global-batch 64 --no-cpu-baseline > gpurun_out/r02_bench_c5_b64_final.json 2> gpurun_out/r02_bench_c5_b64_final.err
# ncu --set full of the headline's dominant kernel and of the EfficientDet stream
A="--only c2 --only-step --no-graph --steps 3 --warmup 3 --repeats 1"
python bench.py $A > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'yolo_loss_ignore_lean_kernel|yolo_loss_scan_kernel|yolo_loss_finalize_kernel|yolo_scatter_targets_kernel|fill_zero_multi' -s 15 -c 5 -o gpurun_out/r02_prof_c2_step -f python bench.py $A > gpurun_out/ncu_full_c2.log 2>&1
A="--only c3 --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
python bench.py $A > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:effdet_stream_kernel -s 3 -c 1 -o gpurun_out/r02_prof_stream_final -f python bench.py $A > gpurun_out/ncu_full_c3.log 2>&1
echo done
A="--only c4 --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
python bench.py $A > gpurun_out/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'effdet_filter_kernel|effdet_nms_finalize_kernel' -s 6 -c 2 -o gpurun_out/r02_prof_c4_final -f python bench.py $A > gpurun_out/ncu_full_c4.log 2>&1
A="--only c1 --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
python bench.py $A > gpurun_out/plain_c1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'yolo_nms_finalize_kernel|yolo_decode_filter_kernel|yolo_classes_kernel' -s 9 -c 3 -o gpurun_out/r02_prof_c1_final -f python bench.py $A > gpurun_out/ncu_full_c1.log 2>&1
python scripts/run_one.py c1_b1 6 > gpurun_out/plain_c1b1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'yolo_nms_finalize_kernel' -s 3 -c 1 -o gpurun_out/r02_prof_c1b1_final -f python scripts/run_one.py c1_b1 6 > gpurun_out/ncu_full_c1b1.log 2>&1
echo done2
