#!/bin/bash
# randomised parity campaigns with new seeds on the final kernels (bounded by wall time)
cd $GRAFT_REPO_ROOT
{ echo "== decode_nms_campaign (first seed 910000)"; python tests/stress/decode_nms_campaign.py 100000 910000 150 2>&1 | tail -4
  echo "== effdet_campaign (first seed 920000)"; python tests/stress/effdet_campaign.py 100000 920000 150 2>&1 | tail -4
  echo "== ignore_mask_campaign (first seed 930000)"; python tests/stress/ignore_mask_campaign.py 100000 930000 100 2>&1 | tail -4; } > gpurun_out/r02_stress_final.txt 2>&1
cat gpurun_out/r02_stress_final.txt
