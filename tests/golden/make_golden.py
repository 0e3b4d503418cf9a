"""Generates the golden fixtures under tests/golden/ from the NumPy oracle.

TensorFlow cannot be installed in the build container, so the reference itself cannot produce vectors; these are
the oracle's outputs (plus the hand-derived known answers of SURVEY.md §8c) frozen so that (a) the oracle cannot
drift silently and (b) the GPU box can check the CUDA path against committed numbers without running the oracle.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import tfmv_b200  # noqa: E402,F401
from oracle import effdet as oe  # noqa: E402
from oracle import yolo as oy  # noqa: E402
from tfmv_b200 import synth  # noqa: E402

F = np.float32
OUT = os.path.dirname(os.path.abspath(__file__))


def yolo_fixture():
    rng = np.random.default_rng(4242)
    image, batch = 64, 2
    anc = (synth.yolo_anchors().astype(F) * F(image / 416.0)).astype(F)
    heads = synth.yolo_heads_trained_like(rng, batch, image, objects=6, dups=3)
    heads = [h + rng.standard_normal(h.shape).astype(F) * F(0.3) for h in heads]
    d = dict(anchors=anc, image=np.int32(image), h0=heads[0], h1=heads[1], h2=heads[2])
    for b in range(batch):
        r = oy.get_nms_boxes_ex(*[h[b:b + 1] for h in heads], anc, (image, image), 80, 0.5, 0.3, 0.5, "diou")
        for k in ("boxes", "classes_id", "scores", "confidence", "selected", "cand_anchor"):
            d["nms%d_%s" % (b, k)] = r[k]
    boxes, classes, off = synth.gt_batch(rng, batch, (image, image), max_boxes=12)
    per = [oy.get_targets(boxes[off[b]:off[b + 1]], classes[off[b]:off[b + 1]], anc / F(image), (image, image), 80) for b in range(batch)]
    y_true = [np.stack([p[l] for p in per], 0) for l in range(3)]
    d.update(gt_boxes=boxes, gt_classes=classes, gt_offsets=off)
    for l in range(3):
        nz = np.argwhere(y_true[l] != 0)
        d["t%d_idx" % l] = nz.astype(np.int32)
        d["t%d_val" % l] = y_true[l][tuple(nz.T)]
    for it in ("iou", "ciou"):
        loss, parts, ign = oy.get_loss(y_true, heads, (image, image), anc / F(image), 0.5, it, return_ignore=True)
        d["loss_" + it] = loss
        d["parts_" + it] = parts
        d["ignore_" + it] = np.packbits(ign, axis=1)
    np.savez_compressed(os.path.join(OUT, "yolo_64.npz"), **d)


def effdet_fixture():
    rng = np.random.default_rng(2424)
    cfg = dict(min_level=3, max_level=5, image_size=(64, 64), num_scales=3,
               aspect_ratios=[(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)], anchor_scale=4.0)
    a = oe.Anchors(cfg["min_level"], cfg["max_level"], cfg["image_size"], cfg["num_scales"], cfg["aspect_ratios"], cfg["anchor_scale"])
    C, batch = 11, 2
    boxes, classes, off = synth.gt_batch(rng, batch, (64, 64), max_boxes=6, order="yxyx", classes=10)
    classes = (classes + 1).astype(np.int32)
    d = dict(gt_boxes=boxes, gt_classes=classes, gt_offsets=off, C=np.int32(C))
    per = [a.generate_targets(boxes[off[b]:off[b + 1]], classes[off[b]:off[b + 1]], C) for b in range(batch)]
    L = len(a.boxes)
    tb = [np.stack([p[0][l] for p in per], 0) for l in range(L)]
    tc = [np.stack([p[1][l] for p in per], 0) for l in range(L)]
    tm = [np.stack([p[2][l] for p in per], 0) for l in range(L)]
    pb = [(rng.standard_normal(t.shape).astype(F) * F(0.25)) for t in tb]
    pc = [rng.standard_normal(t.shape).astype(F) for t in tc]
    dec = a.convert_outputs_boxes(pb)
    loss, parts, npos = oe.get_loss(tb, tc, tm, pb, pc, return_parts=True)
    d.update(loss=loss, parts=parts, npos=npos)
    for l in range(L):
        d["anchors%d" % l] = a.boxes[l]
        d["tb%d" % l], d["tcid%d" % l], d["tm%d" % l] = tb[l], np.argmax(tc[l], -1).astype(np.int8), tm[l]
        d["tcsum%d" % l] = tc[l].sum(-1).astype(np.int8)
        d["pb%d" % l], d["pc%d" % l], d["dec%d" % l] = pb[l], pc[l], dec[l]
    for b in range(batch):
        r = a.convert_outputs_one_ex(b, dec, pc)
        for k in ("boxes", "classes_id", "scores", "selected", "cand_anchor"):
            d["post%d_%s" % (b, k)] = r[k]
    np.savez_compressed(os.path.join(OUT, "effdet_64.npz"), **d)


def kat_fixture():
    # hand-derived known answers for the literal inputs the reference ships
    np.savez(os.path.join(OUT, "kat.npz"),
             b1=np.array([[10, 10, 30, 30]], F), b2=np.array([[20, 20, 40, 40]], F),
             iou=F(100.0 / 700.0), eff_diou=F(100.0 / 700.0 - 200.0 / 1800.0), eff_giou=F(100.0 / 700.0 - 200.0 / 900.0),
             yolo_diou=F(100.0 / 700.0 - (1.0 / 9.0) ** 0.6), yolo_ciou=F(100.0 / 700.0 - 1.0 / 9.0))


if __name__ == "__main__":
    yolo_fixture()
    effdet_fixture()
    kat_fixture()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
