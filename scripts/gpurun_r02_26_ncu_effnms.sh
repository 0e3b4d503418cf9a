#!/bin/bash
# ncu --set full of the EfficientDet NMS kernel at D0 batch 128 (c3)
cd $GRAFT_REPO_ROOT
A="--only c3 --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
python bench.py $A > gpurun_out/plain_c3.log 2>&1 || { echo plain failed; tail gpurun_out/plain_c3.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:effdet_nms_finalize_kernel -s 3 -c 1 -o gpurun_out/r02_prof_effnms_d0 -f python bench.py $A > gpurun_out/ncu_full_effnms.log 2>&1
echo "ncu rc=$?"
