"""Randomised campaign for the target paths (run by hand on a GPU box; not collected by pytest).

    python tests/stress/targets_campaign.py [cases] [first_seed]

(1) GetLossFromBoxes (sparse-target fusion) against GetTargetsBatch + GetLoss on the GPU: loss parts equal to fp64
summation order and ignore masks bit-identical, with many boxes per image (several hundred), forced collisions,
images without boxes, out-of-range classes.  (2) TargetBuffers: a random sequence of ground-truth sets written into
the same persistent tensors must equal a fresh dense assignment after every step, bit for bit.  The dense GPU path
itself is checked against the oracle by the pytest suite."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
F = np.float32


def random_gt(rng, batch, image, max_boxes):
    from tfmv_b200 import synth
    boxes, classes, off = synth.gt_batch(rng, batch, (image, image), max_boxes=max_boxes)
    n = len(boxes)
    if n > 4:  # collisions: copies of other boxes (same image or not), pairs and triples
        k = rng.integers(0, n, n // 5)
        boxes[k] = boxes[rng.integers(0, n, len(k))]
    classes = classes.astype(np.int32)
    classes[rng.random(n) < 0.02] = 300
    if batch > 1 and rng.random() < 0.5:  # empty image
        b = int(rng.integers(0, batch))
        keep = np.ones(n, bool); keep[off[b]:off[b + 1]] = False
        cnt = np.diff(off); cnt[b] = 0
        boxes, classes, off = boxes[keep], classes[keep], np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    return boxes, classes, off


def main():
    import torch
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.datasets.coco_dataset import DataGenerator, TargetBuffers
    from tfmv_b200.ai_models.utils.tf_yolo_utils import GetLossFromBoxes, _loss_call
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
    dev = torch.device("cuda:0")
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    bad = 0
    bufs = {}
    for seed in range(first, first + n):
        rng = np.random.default_rng(seed)
        image = int(rng.choice([96, 160, 256, 416]))
        batch = int(rng.integers(1, 6))
        normalised = bool(rng.random() < 0.5)
        iou_type = str(rng.choice(["iou", "diou", "ciou"]))
        anc = (synth.yolo_anchors().astype(F) * F(image / 608.0)).astype(F)
        tanc = anc / F(image) if normalised else anc
        boxes, classes, off = random_gt(rng, batch, image, int(rng.choice([5, 60, 400])))
        heads = [d(h) for h in synth.yolo_heads(rng, batch, image)]
        gen = DataGenerator(80, tanc, (image, image))
        dense = gen.GetTargetsBatch(d(classes), d(boxes), d(off))
        n_img = sum(t.shape[1] * t.shape[2] * 3 for t in dense)
        ign_d = torch.full((batch, n_img), 7, dtype=torch.uint8, device=dev)
        ign_s = torch.full((batch, n_img), 9, dtype=torch.uint8, device=dev)
        _, pd = _loss_call(dense, heads, (image, image), anc, 0.5, iou_type, 0, return_parts=True, ignore_out=ign_d)
        _, ps = GetLossFromBoxes(d(classes), d(boxes), d(off), heads, (image, image), anc, 80, 0.5, iou_type, target_anchors=tanc,
                                 return_parts=True, ignore_out=ign_s)
        ok = torch.equal(ign_d, ign_s) and np.allclose(ps.cpu().numpy(), pd.cpu().numpy(), rtol=1e-6, atol=1e-9)
        # persistent buffers: one buffer per (image, batch, normalised) reused across the cases that share the key
        key = (image, batch, normalised)
        buf = bufs.setdefault(key, TargetBuffers())
        got = gen.GetTargetsBatch(d(classes), d(boxes), d(off), buffers=buf)
        ok = ok and all(torch.equal(a, b) for a, b in zip(got, dense))
        if not ok:
            bad += 1
            print("MISMATCH seed %d: image %d batch %d boxes %d" % (seed, image, batch, len(boxes)), flush=True)
    print("cases %d  failing cases %d  (persistent buffers reused across %d keys)" % (n, bad, len(bufs)))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
