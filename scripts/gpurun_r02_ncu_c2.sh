#!/bin/bash
# ncu --set full of the c2 step's kernels (lean ignore pass, scan, finalize with the exact pass, scatter)
cd $GRAFT_REPO_ROOT
A="--only c2 --only-step --no-graph --steps 3 --warmup 3 --repeats 1"
python bench.py $A > gpurun_out/plain_c2.log 2>&1 || { echo plain failed; tail gpurun_out/plain_c2.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'yolo_loss_ignore_lean_kernel|yolo_loss_scan_kernel|yolo_loss_finalize_kernel|yolo_scatter_targets_kernel|fill_zero_multi' -s 16 -c 5 -o gpurun_out/r02_prof_c2_step -f python bench.py $A > gpurun_out/ncu_full_c2.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
