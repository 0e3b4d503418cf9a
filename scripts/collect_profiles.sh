#!/bin/bash
# copies the artifacts of scripts/gpurun_r02_final_n1.sh from gpurun_out/ into profiles/ and renders the ncu summaries
cd "$(dirname "$0")/.."
G=gpurun_out
for c in c1 c1_b1 c2 c3 c4 c5 c5_b64; do cp $G/r02_launches_final_$c.csv profiles/; done
cp $G/r02_kernels_summary.txt $G/r02_kernels.json profiles/
cp $G/r02_bench_n1_final.json profiles/r02_bench_n1.json
cp $G/r02_bench_reference.json profiles/
cp $G/r02_bench_c5_b64_final.json profiles/r02_bench_c5_b64.json
cp $G/r02_pytest_final.log profiles/r02_pytest_gpu.log
cp $G/r02_smoke_final.log profiles/r02_smoke.log
python scripts/ncu_summary.py $G/r02_prof_c2_step.ncu-rep yolo_loss_ignore_lean 30 > profiles/r02_ncu_ignore_lean.txt
python scripts/ncu_summary.py $G/r02_prof_c2_step.ncu-rep 'yolo_loss_scan|yolo_loss_finalize|yolo_scatter|fill_zero' 0 | grep -v "^total samples" > profiles/r02_ncu_c2_step_others.txt
python scripts/ncu_summary.py $G/r02_prof_stream_final.ncu-rep effdet_stream 30 > profiles/r02_ncu_stream_final.txt
python scripts/ncu_summary.py $G/r02_prof_c1_final.ncu-rep 'yolo_' 12 > profiles/r02_ncu_c1_final.txt
python scripts/ncu_summary.py $G/r02_prof_c4_final.ncu-rep 'effdet_' 12 > profiles/r02_ncu_c4_final.txt
python scripts/ncu_summary.py $G/r02_prof_c1b1_final.ncu-rep 'yolo_nms' 0 | grep -v "^total samples" > profiles/r02_ncu_c1_b1_final.txt
wc -l profiles/r02_ncu_*.txt
