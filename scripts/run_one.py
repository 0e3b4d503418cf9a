#!/usr/bin/env python
"""Small driver for ncu captures: runs one entry point a few times launch by launch.
  python scripts/run_one.py c1_b1 | c5_b64 | c3 | c4"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import tfmv_b200  # noqa: E402,F401
from tfmv_b200 import synth  # noqa: E402
from tfmv_b200.ai_models.utils import tf_yolo_utils as tyu  # noqa: E402

what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
anc = synth.yolo_anchors().astype(np.float32)
if what == "c1_b1":
    heads = [torch.randn((1, s, s, 255), device=dev, generator=g) for s in synth.yolo_grids(416)]
    for _ in range(n):
        r = tyu.GetNMSBoxesBatch(*heads, anc, (416, 416), 80, 0.5, 0.3, 0.5, "iou")
    torch.cuda.synchronize()
    print("kept", int(r["count"][0]))
elif what == "c5_b64":
    from tfmv_b200.ai_models.datasets.coco_dataset import DataGenerator
    B = 64
    heads = [torch.randn((B, s, s, 255), device=dev, generator=g) for s in synth.yolo_grids(608)]
    rng = np.random.default_rng(5)
    boxes, classes, off = synth.gt_batch(rng, B, (608, 608), max_boxes=100)
    gen = DataGenerator(80, anc, (608, 608))
    to = lambda a: torch.from_numpy(a).to(dev)
    y_true = gen.GetTargetsBatch(to(classes), to(boxes), to(off))
    for _ in range(n):
        loss, r = tyu.LossAndNMSBoxesBatch(y_true, heads, (608, 608), anc, 80, 0.5, "ciou", 0.5, 0.3, 0.5, "diou")
    torch.cuda.synchronize()
    print("loss", float(loss))
