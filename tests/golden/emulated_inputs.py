"""Seeded inputs shared by tests/golden/make_golden_emulated.py (which runs the reference on them) and the tests (which
run the oracle / the CUDA path on them), for cases whose inputs are too large to store next to the outputs."""
import numpy as np

F = np.float32
COCO_ANCHORS = np.array([116, 90, 156, 198, 373, 326, 30, 61, 62, 45, 59, 119, 10, 13, 16, 30, 33, 23], F).reshape(3, 3, 2)


def yolo_416_heads(seed=20261018 + 416):
    """YOLOv3 416x416, 80 classes, one image, every logit ~ N(0,1) (BASELINE config 1): ~5.3 k candidates, cap 500 reached."""
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((1, g, g, 255), dtype=F) for g in (13, 26, 52)]


D0 = dict(min_level=3, max_level=7, image_size=(512, 512), num_scales=3, aspect_ratios=[(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)],
          anchor_scale=4.0)


def effdet_d0_heads(level_shapes, seed=20261018 + 512, classes=81):
    """EfficientDet-D0 512x512, one image (BASELINE config 3 per image): box outputs ~ 0.25 N(0,1), class logits ~ N(0,1).
    level_shapes: [(H, W, A)] of the pyramid."""
    rng = np.random.default_rng(seed)
    rel = [(rng.standard_normal((1, h, w, a, 4), dtype=F) * F(0.25)) for h, w, a in level_shapes]
    cls = [rng.standard_normal((1, h, w, a, classes), dtype=F) for h, w, a in level_shapes]
    return rel, cls
