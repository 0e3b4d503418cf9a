"""Synthetic inputs of the shapes BASELINE.md names (host-side NumPy; used by tests and bench.py).

All tensors fp32.  RNG: np.random.default_rng(20261018 + config_id) at the call sites.
"""
import numpy as np

F = np.float32
COCO_ANCHORS = np.array([10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326], dtype=np.int64)


def yolo_anchors():
    """(3,3,2) pixels, layer 0 = stride-32 anchors (LoadAnchors order, load_object_detection_data.py:58-67)."""
    return COCO_ANCHORS.reshape(3, 3, 2)[[2, 1, 0]].copy()


def yolo_grids(image):
    return [image // 32, image // 16, image // 8]


def yolo_heads(rng, batch, image=416, classes=80, anchors=3):
    """'random-init' heads: every logit ~ N(0,1); list of 3 arrays (B,H,W,A*(5+C))."""
    return [rng.standard_normal((batch, g, g, anchors * (5 + classes)), dtype=F) for g in yolo_grids(image)]


def yolo_heads_trained_like(rng, batch, image=416, classes=80, anchors=3, objects=50, dups=5):
    """conf ~ N(-4,1.5) background with `objects` planted detections x `dups` jittered duplicates per image."""
    heads = []
    for g in yolo_grids(image):
        h = rng.standard_normal((batch, g, g, anchors, 5 + classes), dtype=F)
        h[..., 4] = h[..., 4] * F(1.5) - F(4.0)
        h[..., 5:] = h[..., 5:] - F(3.0)
        heads.append(h)
    for b in range(batch):
        for _ in range(objects):
            l = int(rng.integers(0, 3))
            g = heads[l].shape[1]
            cy, cx = int(rng.integers(0, g)), int(rng.integers(0, g))
            cls = int(rng.integers(0, classes))
            twh = rng.normal(0, 0.5, size=2).astype(F)
            for _ in range(dups):
                y = min(max(cy + int(rng.integers(-1, 2)), 0), g - 1)
                x = min(max(cx + int(rng.integers(-1, 2)), 0), g - 1)
                a = int(rng.integers(0, anchors))
                rec = heads[l][b, y, x, a]
                rec[0:2] = rng.normal(0, 1, size=2)
                rec[2:4] = twh + rng.normal(0, 0.05, size=2)
                rec[4] = F(3.0 + rng.random())
                rec[5 + cls] = F(2.0 + 2.0 * rng.random())
    return [h.reshape(h.shape[0], h.shape[1], h.shape[2], -1) for h in heads]


def gt_boxes(rng, image_wh, max_boxes=100, classes=80, order="xyxy", min_boxes=1):
    """n ~ U{min..max}; centres U(.05,.95)*image; w,h log-uniform in [.02,.6]*image; clipped to the image."""
    n = int(rng.integers(min_boxes, max_boxes + 1))
    iw, ih = float(image_wh[0]), float(image_wh[1])
    cx = rng.uniform(0.05, 0.95, n) * iw
    cy = rng.uniform(0.05, 0.95, n) * ih
    w = np.exp(rng.uniform(np.log(0.02), np.log(0.6), n)) * iw
    h = np.exp(rng.uniform(np.log(0.02), np.log(0.6), n)) * ih
    x1, x2 = np.clip(cx - w / 2, 0, iw), np.clip(cx + w / 2, 0, iw)
    y1, y2 = np.clip(cy - h / 2, 0, ih), np.clip(cy + h / 2, 0, ih)
    cls = rng.integers(0, classes, n).astype(np.int32)
    if order == "xyxy":
        return np.stack([x1, y1, x2, y2], -1).astype(F), cls
    return np.stack([y1, x1, y2, x2], -1).astype(F), cls


def gt_batch(rng, batch, image_wh, **kw):
    """Ragged GT for a batch: (boxes [total,4], classes [total], offsets [B+1] int32)."""
    bs, cs, off = [], [], [0]
    for _ in range(batch):
        b, c = gt_boxes(rng, image_wh, **kw)
        bs.append(b)
        cs.append(c)
        off.append(off[-1] + b.shape[0])
    return np.concatenate(bs, 0), np.concatenate(cs, 0), np.asarray(off, dtype=np.int32)


EFFDET_CONFIGS = {
    # utils/global_params.py:144-197 defaults; D7 overrides anchor_scale=5.0, image_size=1536 (:110-124)
    "d0": dict(image_size=(512, 512), min_level=3, max_level=7, num_scales=3,
               aspect_ratios=[(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)], anchor_scale=4.0, num_classes=81),
    "d7": dict(image_size=(1536, 1536), min_level=3, max_level=7, num_scales=3,
               aspect_ratios=[(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)], anchor_scale=5.0, num_classes=81),
}


def effdet_level_sizes(image_size, min_level=3, max_level=7):
    fs = (int(image_size[0]), int(image_size[1]))
    sizes = [fs]
    for _ in range(1, max_level + 1):
        fs = ((fs[0] - 1) // 2 + 1, (fs[1] - 1) // 2 + 1)
        sizes.append(fs)
    return sizes[min_level:max_level + 1]


def effdet_heads(rng, batch, image_size=(512, 512), num_classes=81, anchors=9, min_level=3, max_level=7):
    """class logits ~ N(0,1) (B,H,W,9,81); box regressions ~ N(0,0.25) (B,H,W,9,4); one pair per level."""
    boxes, classes = [], []
    for (h, w) in effdet_level_sizes(image_size, min_level, max_level):
        classes.append(rng.standard_normal((batch, h, w, anchors, num_classes), dtype=F))
        boxes.append(rng.standard_normal((batch, h, w, anchors, 4), dtype=F) * F(0.25))
    return boxes, classes
