#!/bin/bash
# ncu --set full of the c1 kernels: NMS CTA at batch 1, filter / NMS / class rows at batch 256
cd $GRAFT_REPO_ROOT
python scripts/run_one.py c1_b1 6 > gpurun_out/plain_c1b1.log 2>&1 || { echo plain failed; tail gpurun_out/plain_c1b1.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'yolo_nms_finalize_kernel|yolo_decode_filter_kernel|yolo_classes_kernel' -s 9 -c 3 -o gpurun_out/r02_prof_c1b1_v2 -f python scripts/run_one.py c1_b1 6 > gpurun_out/ncu_full_c1b1.log 2>&1
echo "ncu b1 rc=$?"
A="--only c1 --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
python bench.py $A > gpurun_out/plain_c1.log 2>&1 || { echo plain c1 failed; tail gpurun_out/plain_c1.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'yolo_nms_finalize_kernel|yolo_decode_filter_kernel|yolo_classes_kernel' -s 9 -c 3 -o gpurun_out/r02_prof_c1b256_v2 -f python bench.py $A > gpurun_out/ncu_full_c1.log 2>&1
echo "ncu b256 rc=$?"; ls -la gpurun_out/r02_prof_c1*.ncu-rep
