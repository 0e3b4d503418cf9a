#!/bin/bash
# headline step after a yolo_loss change: parity tests + step time (three runs)
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_yolo_loss.py tests/test_golden.py tests/test_gpu_fullsize_properties.py tests/test_gpu_reentrancy.py tests/test_gpu_reference_emulated.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2 3; do python bench.py --only c2 --only-step --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2', d['ms_per_step'], round(d['value']))"; done
python bench.py --only c5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5', d['configs']['c5']['ms_per_step'])"
python tests/stress/ignore_mask_campaign.py 100000 940000 40 2>&1 | tail -2
