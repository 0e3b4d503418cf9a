#!/usr/bin/env python
"""Device time of GetNMSBoxesBatch alone: python scripts/time_nms_batch.py image batch [metric]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import tfmv_b200  # noqa: E402,F401
from tfmv_b200 import synth  # noqa: E402
from tfmv_b200.ai_models.utils import tf_yolo_utils as tyu  # noqa: E402

image, batch = int(sys.argv[1]), int(sys.argv[2])
metric = sys.argv[3] if len(sys.argv) > 3 else "diou"
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
anc = synth.yolo_anchors().astype(np.float32)
heads = [torch.randn((batch, s, s, 255), device=dev, generator=g) for s in synth.yolo_grids(image)]
fn = lambda: tyu.GetNMSBoxesBatch(*heads, anc, (image, image), 80, 0.5, 0.3, 0.5, metric)
for _ in range(5):
    fn()
ts = []
for rep in range(7):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(20):
        fn()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) / 20 * 1e3)
print("image %d batch %d %s: %.1f us per call (min %.1f)" % (image, batch, metric, float(np.median(ts)), min(ts)))
