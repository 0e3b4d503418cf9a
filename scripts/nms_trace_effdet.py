#!/usr/bin/env python
"""Phase times of the EfficientDet per-image NMS CTA (block 0) from a -DNMS_TRACE build.
  python scripts/nms_trace_effdet.py d0|d7 [batch]"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import tfmv_b200  # noqa: E402,F401
from tfmv_b200 import _lib, synth  # noqa: E402
from tfmv_b200.ai_models.efficientnet.utils.anchors import Anchors  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "d0"
size = {"d0": 512, "d7": 1536}[name]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else (8 if name == "d0" else 2)
dev = torch.device("cuda:0")
rng = np.random.default_rng(3)
a = Anchors(3, 7, (size, size), 3, [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)], 4.0)
boxes, classes = synth.effdet_heads(rng, batch, (size, size))
rel = [torch.from_numpy(b).to(dev) for b in boxes]
cls = [torch.from_numpy(c).to(dev) for c in classes]
lib = ctypes.CDLL(_lib.load()._name)
names = {13: "entry", 0: "count read", 1: "G1 range pass", 2: "G2 histogram pass", 3: "G3 bin scan", 4: "G4 gather pass",
         5: "window ordered", 6: "chunk boxes loaded", 20: "tiles consumed", 11: "NMS done", 12: "outputs written"}
order = [13, 0, 1, 2, 3, 4, 5, 6, 20, 11, 12]
acc = {k: [] for k in order}
for it in range(6):
    dec, r = a.decode_and_postprocess(rel, cls)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 192)()
    assert lib.b200_debug_nms_trace_effdet(buf) == 0
    if it >= 2:
        for k in order:
            acc[k].append(buf[k])
print(name, "batch", batch, "kept", int(r["count"][0]), "tiles in the last chunk", buf[21], "windows (cumulative)", buf[22])
prev = None
for k in order:
    t = np.array(acc[k], dtype=np.int64)
    if prev is not None:
        d = (t - prev) / 1965.0
        print("%-24s %7.2f us  (min %.2f max %.2f)" % (names[k], float(np.median(d)), d.min(), d.max()))
    prev = t
print("   tile phases of the last chunk (thread 0, arrival to arrival), us: vs kept %.2f | intra-tile mask %.2f | sweep %.2f | emit %.2f" % tuple(buf[k] / 1965.0 for k in (24, 25, 26, 27)))
print("   pair tests (cumulative over %d runs): %d, of which full metric: %d" % (6, buf[30], buf[31]))
print("total %.2f us" % float(np.median((np.array(acc[12]) - np.array(acc[13])) / 1965.0)))
