"""Randomised campaign for GetNMSBoxes (run by hand on a GPU box; not collected by pytest).

    python tests/stress/decode_nms_campaign.py [cases] [first_seed] [seconds]

Random image sizes, batch sizes, thresholds and metrics; logits of several spreads (saturating sigmoids, exp overflow
-> box dropped), conf logits planted right around logit(conf_thr), class logits with near-ties inside and outside the
guard band, duplicated records (exact score ties).  Indices, class ids, boxes, scores, classes and confidences must
equal the oracle's bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
F = np.float32


def one_case(seed, dev):
    import torch
    from oracle import yolo as oy
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.utils.tf_yolo_utils import GetNMSBoxesBatch
    rng = np.random.default_rng(seed)
    image = int(rng.choice([64, 96, 128, 160]))
    batch = int(rng.integers(1, 5))
    conf_thr = float(rng.choice([0.5, 0.3, 0.7, 0.05, 0.95]))
    score_thr = float(rng.choice([0.3, 0.2, 0.5, 0.9]))
    iou_thr = float(rng.choice([0.5, 0.45, 0.3]))
    iou_type = str(rng.choice(["iou", "diou", "ciou"]))
    sigma = float(rng.choice([1.0, 1.0, 3.0, 8.0, 30.0]))
    classes_num = int(rng.choice([80, 80, 20, 3]))
    heads = [(rng.standard_normal((batch, g, g, 3, 5 + classes_num)) * sigma).astype(F) for g in synth.yolo_grids(image)]
    t = np.log(conf_thr / (1 - conf_thr))
    for h in heads:
        flat = h.reshape(-1, 5 + classes_num)
        n = flat.shape[0]
        k = rng.integers(0, n, max(1, n // 10))
        flat[k, 4] = (t + rng.normal(0, 3e-6, len(k)) * rng.choice([1.0, 10.0, 1000.0], len(k))).astype(F)  # conf at the threshold
        k = rng.integers(0, n, max(1, n // 10))
        c1, c2 = rng.integers(0, classes_num, len(k)), rng.integers(0, classes_num, len(k))
        top = np.abs(flat[k]).max(1) + rng.uniform(0.1, 2.0, len(k)).astype(F)
        flat[k, 5 + c1] = top
        flat[k, 5 + c2] = top - rng.choice([0.0, 1e-7, 1e-5, 5e-3, 0.02], len(k)).astype(F)             # near ties of the class maximum
        k = rng.integers(0, n - 1, max(1, n // 20))
        flat[k + 1] = flat[k]                                                                             # duplicated records
        k = rng.integers(0, n, max(1, n // 50))
        flat[k, 2] = rng.choice([95.0, -95.0, 88.0, -40.0], len(k)).astype(F)                             # exp overflow / tiny boxes
    heads = [h.reshape(batch, h.shape[1], h.shape[2], -1) for h in heads]
    anc = synth.yolo_anchors()
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    r = GetNMSBoxesBatch(*[d(h) for h in heads], anc, (image, image), classes_num, conf_thr, score_thr, iou_thr, iou_type,
                         with_indices=True)
    r = {k: v.cpu().numpy() for k, v in r.items()}
    from test_gpu_yolo_decode import _check_image
    emitted = 0
    for b in range(batch):
        want = oy.get_nms_boxes_ex(*[h[b:b + 1] for h in heads], anc, (image, image), classes_num, conf_thr, score_thr, iou_thr, iou_type)
        _check_image(r, b, want, classes_num)
        emitted += want["selected"].shape[0]
    return emitted


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
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
