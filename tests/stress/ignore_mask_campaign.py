"""Randomised campaign for the ignore mask of GetLoss (run by hand on a GPU box; not collected by pytest).

    python tests/stress/ignore_mask_campaign.py [cases] [first_seed] [seconds]

Every case draws an image size, a threshold >= 0.5 (the regime of the decode-free and approximate-IoU rejects,
DESIGN.md §6), a metric, a logit spread and a ground-truth set that includes very small boxes, plants predictions
whose IoU with their target is within a few per cent of the threshold (concentric boxes of area ratio thr*(1+eps)),
and compares the GPU's ignore mask with the oracle's bit for bit, for both the dense and the sparse-target entry."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
F = np.float32


def one_case(seed, dev):
    import torch
    from oracle import yolo as oy
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.utils.tf_yolo_utils import _loss_call
    rng = np.random.default_rng(seed)
    image = int(rng.choice([96, 128, 256]))
    batch = int(rng.integers(1, 4))
    thr = float(rng.choice([0.5, 0.5, 0.55, 0.6, 0.7, 0.9]))
    iou_type = str(rng.choice(["iou", "diou", "ciou"]))
    sigma = float(rng.choice([1.0, 2.0, 3.0, 5.0]))
    anc = (synth.yolo_anchors().astype(F) * F(image / 608.0)).astype(F)
    n_max = int(rng.integers(1, 60))
    boxes, classes, off = synth.gt_batch(rng, batch, (image, image), max_boxes=n_max)
    # a share of very small boxes (down to about one pixel) to exercise the extent floors
    tiny = rng.random(len(boxes)) < 0.3
    cx, cy = (boxes[:, 0] + boxes[:, 2]) / 2, (boxes[:, 1] + boxes[:, 3]) / 2
    w = np.where(tiny, rng.uniform(1.0, 6.0, len(boxes)), boxes[:, 2] - boxes[:, 0])
    h = np.where(tiny, rng.uniform(1.0, 6.0, len(boxes)), boxes[:, 3] - boxes[:, 1])
    boxes = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], -1).clip(0, image).astype(F)
    per = [oy.get_targets(boxes[off[b]:off[b + 1]], classes[off[b]:off[b + 1]], anc / F(image), (image, image), 80) for b in range(batch)]
    y_true = [np.stack([p[l] for p in per], 0) for l in range(3)]
    y_pred = [h_ * F(sigma) for h_ in synth.yolo_heads(rng, batch, image)]
    planted = 0
    for l in range(3):
        yt = y_true[l]
        yp = y_pred[l].reshape(yt.shape)
        b, yy, xx, aa = np.nonzero(yt[..., 4] > 0)
        g = yt.shape[1]
        for k in range(len(b)):
            t = yt[b[k], yy[k], xx[k], aa[k]]
            eps = rng.uniform(-0.03, 0.03)
            ratio = min(thr * (1.0 + eps), 1.0)             # concentric boxes: IoU = area ratio
            s = np.sqrt(ratio) if k % 2 == 0 else 1.0 / np.sqrt(ratio)
            for a2 in range(3):
                if rng.random() < 0.5:
                    continue
                fx, fy = t[0] * g - xx[k], t[1] * g - yy[k]
                fx, fy = min(max(fx, 1e-3), 1 - 1e-3), min(max(fy, 1e-3), 1 - 1e-3)
                yp[b[k], yy[k], xx[k], a2, 0] = np.log(fx / (1 - fx))
                yp[b[k], yy[k], xx[k], a2, 1] = np.log(fy / (1 - fy))
                yp[b[k], yy[k], xx[k], a2, 2:4] = np.log(np.maximum(t[2:4] * image * s, 1e-6) / anc[l][a2])
                planted += 1
    want, _, want_ign = oy.get_loss(y_true, y_pred, (image, image), anc, thr, iou_type, return_ignore=True)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    ign = torch.full(want_ign.shape, 7, dtype=torch.uint8, device=dev)
    _loss_call([d(t) for t in y_true], [d(t) for t in y_pred], (image, image), anc, thr, iou_type, 0, ignore_out=ign)
    got = ign.cpu().numpy()
    bad = int((got != want_ign).sum())
    return bad, int((want_ign == 0).sum()), planted, (image, batch, thr, iou_type, sigma)


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    budget = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0   # optional wall-clock budget in seconds
    import time
    t_start, done = time.time(), 0
    dev = torch.device("cuda:0")
    total_bad = zeros = planted = 0
    for seed in range(first, first + n):
        if budget and time.time() - t_start > budget:
            break
        done += 1
        bad, z, pl, cfg = one_case(seed, dev)
        total_bad += bad
        zeros += z
        planted += pl
        if bad:
            print("MISMATCH seed %d: %d bits, case %r" % (seed, bad, cfg), flush=True)
    print("cases %d  planted predictions %d  ignore-mask zeros %d  mismatching bits %d" % (done, planted, zeros, total_bad))
    return 1 if total_bad else 0


if __name__ == "__main__":
    sys.exit(main())
