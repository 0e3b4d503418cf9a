#!/bin/bash
# NMS kernel duration at batch 1 (ncu launch list) for the library that is in the tree
cd $GRAFT_REPO_ROOT
python scripts/run_one.py c1_b1 6 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/b1_tmp.csv python scripts/run_one.py c1_b1 8 > /dev/null 2>&1
for k in yolo_nms_finalize yolo_decode_filter yolo_classes; do echo -n "$k: "; grep "$k" gpurun_out/b1_tmp.csv | grep time_duration | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo; done
