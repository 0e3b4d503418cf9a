#!/usr/bin/env python
"""Phase times of the per-image NMS CTA (block 0) from a -DNMS_TRACE build:
  B200_EXTRA_NVCC_FLAGS=-DNMS_TRACE python tensorflow2-machine-vision_b200/build.py --force
  python scripts/nms_trace.py [batch] [image] [iou|diou|ciou]"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import tfmv_b200  # noqa: E402,F401
from tfmv_b200 import _lib, synth  # noqa: E402
from tfmv_b200.ai_models.utils import tf_yolo_utils as tyu  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
image = int(sys.argv[2]) if len(sys.argv) > 2 else 416
metric = sys.argv[3] if len(sys.argv) > 3 else "iou"
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
anc = synth.yolo_anchors().astype(np.float32)
heads = [torch.randn((batch, s, s, 255), device=dev, generator=g) for s in synth.yolo_grids(image)]
lib = ctypes.CDLL(_lib.load()._name)
names = {13: "entry", 0: "count read, heads cleared", 1: "G1 range pass", 2: "G2 histogram pass", 3: "G3 bin scan", 4: "G4 gather pass",
         5: "window ordered (buckets)", 6: "chunk boxes decoded", 7: "(1) vs kept + table clear", 8: "(2) class split", 9: "(3) buckets resolved",
         10: "(4) survivors emitted", 11: "NMS done", 12: "outputs written"}
order = [13, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12]
acc = {k: [] for k in order}
for it in range(8):
    r = tyu.GetNMSBoxesBatch(*heads, anc, (image, image), 80, 0.5, 0.3, 0.5, metric)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 192)()
    assert lib.b200_debug_nms_trace(buf) == 0
    if it >= 3:
        for k in order:
            acc[k].append(buf[k])
print("kept", int(r["count"][0]), "batch", batch, "image", image)
prev = None
for k in order:
    t = np.array(acc[k], dtype=np.int64)
    if prev is not None:
        d = (t - prev) / 1965.0
        print("%-30s %7.2f us  (min %.2f max %.2f)" % (names[k], float(np.median(d)), d.min(), d.max()))
    prev = t
extra = {16: "order: range", 17: "order: histogram + chains", 18: "order: scan", 19: "order: ranks", 14: "(4) after count barrier", 15: "(4) thread 0 stores done"}
base = {16: 4, 17: 16, 18: 17, 19: 18, 14: 9, 15: 14}
for k in (16, 17, 18, 19, 14, 15):
    print("   %-28s %7.2f us" % (extra[k], (buf[k] - buf[base[k]]) / 1965.0))
for k, nm in enumerate(("(3) pair pass start", "(3) pair pass end", "(3) resolve start", "(3) resolve end")):
    print("   %-22s per warp, us after T8:" % nm, " ".join("%.1f" % ((buf[64 + k * 32 + w] - buf[8]) / 1965.0) for w in range(32)))
tot = (np.array(acc[12]) - np.array(acc[13])) / 1965.0
print("total %.2f us" % float(np.median(tot)))
