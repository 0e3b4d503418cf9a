"""Randomised campaign for the EfficientDet utilities (run by hand on a GPU box; not collected by pytest).

    python tests/stress/effdet_campaign.py [cases] [first_seed] [seconds]

Random anchor configurations (levels, scales, aspect ratios, image sizes incl. rectangular), batch sizes and class
counts.  Post-processing: logits with exact ties between anchors, ties across classes, planted duplicates (heavy
suppression), several metrics; selected indices, class ids, boxes and scores must equal the oracle's bit for bit.
Target assignment: random ground truth with out-of-range classes, thresholds including 0; boxes, one-hot rows (and
class ids) and masks bit-exact.  Loss within 1e-4."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
F = np.float32


def one_case(seed, dev):
    import torch
    from oracle import effdet as oe
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.efficientnet.efficientdet_net_train import get_loss
    from tfmv_b200.ai_models.efficientnet.utils.anchors import Anchors
    rng = np.random.default_rng(seed)
    min_level = int(rng.integers(2, 4))
    max_level = min_level + int(rng.integers(1, 4))
    size = (int(rng.choice([64, 96, 128, 160])), int(rng.choice([64, 96, 128, 192])))
    num_scales = int(rng.integers(1, 4))
    aspects = [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)][:int(rng.integers(1, 4))]
    scale = float(rng.choice([2.0, 3.0, 4.0]))
    args = (min_level, max_level, size, num_scales, aspects, scale)
    if seed % 25 == 0:  # a full EfficientDet-D0 pyramid (49 104 anchors: the multi-CTA pre-selection path of the NMS)
        args = (3, 7, (512, 512), 3, [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)], 4.0)
        size = (512, 512)
    a, o = Anchors(*args), oe.Anchors(*args)
    batch = 1 if seed % 25 == 0 else int(rng.integers(1, 4))
    C = int(rng.choice([3, 21, 81]))
    iou_type = str(rng.choice(["iou", "giou", "diou", "ciou"]))
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    # ---- post-processing ----
    rel = [(rng.standard_normal((batch,) + b.shape, dtype=F) * F(rng.choice([0.1, 0.25, 1.0]))) for b in o.boxes]
    cls = [(rng.standard_normal((batch,) + b.shape[:-1] + (C,), dtype=F) * F(rng.choice([1.0, 3.0]))) for b in o.boxes]
    for l in range(len(cls)):
        flat = cls[l].reshape(-1, C)
        n = flat.shape[0]
        k = rng.integers(0, n, max(1, n // 8))
        flat[k] = flat[rng.integers(0, n, len(k))]            # identical rows: exact score ties between anchors
        k = rng.integers(0, n, max(1, n // 16))
        flat[k, :] = F(0.5)                                    # ties across classes -> background
        rflat = rel[l].reshape(-1, 4)
        if n > 1:                                              # (a 1x1 level with one anchor and one image has no neighbour)
            k = rng.integers(0, n - 1, max(1, n // 6))
            rflat[k + 1] = rflat[k]                            # neighbouring anchors predicting (almost) the same box
    dec_w = o.convert_outputs_boxes(rel)
    dec = a.convert_outputs_boxes([d(r) for r in rel])
    for x, y in zip(dec, dec_w):
        assert np.array_equal(x.cpu().numpy().view(np.uint32), y.view(np.uint32)), "decode"
    r = a.convert_outputs_batch(dec, [d(c) for c in cls], iou_type=iou_type, with_indices=True)
    r = {k: v.cpu().numpy() for k, v in r.items()}
    emitted = 0
    for b in range(batch):
        w = o.convert_outputs_one_ex(b, dec_w, cls, iou_type=iou_type)
        k = int(r["count"][b])
        assert k == w["selected"].shape[0], "count"
        assert r["sel_idx"][b, :k].tolist() == w["selected"].tolist(), "selected"
        assert r["classes_id"][b, :k].tolist() == w["classes_id"].tolist(), "class ids"
        assert np.array_equal(r["boxes"][b, :k].view(np.uint32), w["boxes"].view(np.uint32)), "boxes"
        assert np.array_equal(r["scores"][b, :k].view(np.uint32), w["scores"].view(np.uint32)), "scores"
        emitted += k
    # ---- target assignment + loss ----
    boxes, classes, off = synth.gt_batch(rng, batch, (size[1], size[0]), max_boxes=int(rng.choice([3, 20, 60])), order="yxyx")
    classes = (classes % C + 1).astype(np.int32)   # includes the out-of-range id C
    thr = float(rng.choice([0.5, 0.5, 0.3, 0.0]))
    gb, gc, gm = a.generate_targets_batch(d(boxes), d(classes), d(off), C, iou_threshold=thr)
    ib, ic, im = a.generate_targets_batch(d(boxes), d(classes), d(off), C, iou_threshold=thr, class_index=True)
    tb, tc, tm = [], [], []
    for b in range(batch):
        wb, wc, wm = o.generate_targets(boxes[off[b]:off[b + 1]], classes[off[b]:off[b + 1]], C, iou_threshold=thr)
        tb.append(wb); tc.append(wc); tm.append(wm)
        for l in range(len(wb)):
            assert np.array_equal(gb[l][b].cpu().numpy().view(np.uint32), wb[l].view(np.uint32)), "target boxes"
            assert np.array_equal(gc[l][b].cpu().numpy(), wc[l]), "one-hot"
            assert np.array_equal(gm[l][b].cpu().numpy(), wm[l]), "mask"
            ids = ic[l][b].cpu().numpy()
            oh = np.zeros_like(wc[l]); ok = (ids >= 0) & (ids < C)
            oh[ok, ids[ok]] = 1
            assert np.array_equal(oh, wc[l]) and torch.equal(ib[l][b], gb[l][b]) and torch.equal(im[l][b], gm[l][b]), "class ids"
    L = len(o.boxes)
    stack = lambda xs, l: np.stack([x[l] for x in xs], 0)
    want = oe.get_loss([stack(tb, l) for l in range(L)], [stack(tc, l) for l in range(L)], [stack(tm, l) for l in range(L)], rel, cls)
    got = get_loss(gb, gc, gm, [d(x) for x in rel], [d(x) for x in cls])
    got_i = get_loss(ib, ic, im, [d(x) for x in rel], [d(x) for x in cls])
    assert abs(float(got) - float(want)) <= 1e-4 * abs(float(want)), "loss"
    assert abs(float(got_i) - float(want)) <= 1e-4 * abs(float(want)), "loss (class ids)"
    return emitted


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 12000
    budget = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0   # optional wall-clock budget in seconds
    import time
    t_start, done = time.time(), 0
    dev = torch.device("cuda:0")
    bad = emitted = 0
    for seed in range(first, first + n):
        if budget and time.time() - t_start > budget:
            break
        done += 1
        try:
            emitted += one_case(seed, dev)
        except AssertionError as e:
            bad += 1
            print("MISMATCH seed %d: %s" % (seed, str(e)[:200]), flush=True)
    print("cases %d  emitted boxes %d  failing cases %d" % (done, emitted, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
