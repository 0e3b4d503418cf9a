"""Golden vectors for the serving view's box post-processing, produced by the REFERENCE's own lines.

views/object_detection.py is a Django view and cannot be imported here, but lines 71-85 of `predict` (scale back, clip,
size filter, int cast) are plain NumPy statements on local variables.  This script reads exactly those lines from
/root/reference at generation time (nothing is copied into the repository), dedents them and executes them on seeded
inputs.  The reference dates from NumPy 1.x, where `float32_array * np.int32_scalar` stays float32 (value-based
casting); under this container's NumPy 2 the same expression would promote to float64, so `image_size` and
`image_size_old` are handed over as Python ints, which NumPy 2 treats as weak scalars — the float32 arithmetic of the
reference's era.  Re-run: python tests/golden/make_golden_views.py
"""
import hashlib
import os
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/AIServer/ai_api/views/object_detection.py"
OUT = os.path.join(HERE, "ref_views.npz")
CASES = [((640, 480), (52, 52, 0, 0)), ((333, 777), (0, 0, 119, 119)), ((416, 416), (0, 0, 0, 0)), ((1920, 1080), (91, 91, 0, 0)),
         ((4032, 3024), (52, 52, 0, 0)), ((37, 91), (0, 0, 123, 124))]


def inputs(i, old_wh, n=600):
    rng = np.random.default_rng(20261018 + 100 + i)
    c = rng.uniform(-0.1, 1.1, (n, 2))
    wh = np.exp(rng.uniform(np.log(1e-3), np.log(0.8), (n, 2)))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    boxes[::50] = boxes[1::50]
    boxes[5, 2] = boxes[5, 0] + np.float32(2.0 / old_wh[0])
    return (boxes, rng.integers(0, 80, n).astype(np.int32), rng.random(n).astype(np.float32),
            rng.random((n, 80)).astype(np.float32), rng.random((n, 1)).astype(np.float32))


def main():
    lines = open(REF, encoding="utf-8").read().split("\n")[70:85]   # 1-based lines 71..85
    assert lines[0].lstrip().startswith("y_boxes[:,[0,2]] =") and lines[-1].strip() == "y_boxes = y_boxes.astype(np.int32)", lines
    code = compile(textwrap.dedent("\n".join(lines)), REF + ":71-85", "exec")
    d = {}
    for i, (old_wh, padding) in enumerate(CASES):
        b, cid, sc, cl, cf = inputs(i, old_wh)
        ns = {"np": np, "y_boxes": b.copy(), "y_classes_id": cid, "y_scores": sc, "y_classes": cl, "y_confidence": cf,
              "image_size": [416, 416], "padding": tuple(padding), "image_size_old": [int(old_wh[0]), int(old_wh[1])]}
        exec(code, ns)
        assert ns["y_boxes"].dtype == np.int32
        for k in ("y_boxes", "y_classes_id", "y_scores"):
            d["%d/%s" % (i, k)] = ns[k]
        for k in ("y_classes", "y_confidence"):   # wide random rows: their sha256 is enough
            d["%d/%s_sha" % (i, k)] = np.array(hashlib.sha256(np.ascontiguousarray(ns[k]).tobytes()).hexdigest())
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, len(CASES), "cases", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
