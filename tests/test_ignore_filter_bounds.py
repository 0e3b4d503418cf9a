"""CPU property test of the ignore-mask filter of yolo_loss_ignore_kernel: the three reject conditions
the kernel applies before the exact test — cell rectangle vs GT, logit-sum (area) window, approximate-IoU upper bound —
are restated here in NumPy and checked against the oracle's exact metric on adversarial pairs: a pair whose exact
metric reaches the threshold must never be rejected.  The approximate decode is additionally perturbed by the error
bound assumed for the GPU's MUFU arithmetic (2e-6 per corner)."""
import numpy as np

from oracle import detmath as dm
from oracle import yolo as oy

F = np.float32
CELL_MARGIN, LOG_MARGIN, IOU_EPS, IOU_FLOOR = F(1e-3), F(0.05), F(5e-3), F(1e-2)   # csrc/yolo_loss.cu


def _filters_reject(tx, ty, tw, th, gx, gy, W, H, anc_w, anc_h, img_w, img_h, g, thr, rng):
    """g: GT corners (n,4).  Returns a bool array: True where the kernel's decode-free / approximate rejects fire."""
    n = g.shape[0]
    ga = (g[:, 2] - g[:, 0]) * (g[:, 3] - g[:, 1])
    inv_w, inv_h = F(1) / F(W), F(1) / F(H)
    cx0, cx1 = F(gx) * inv_w - CELL_MARGIN, F(gx + 1) * inv_w + CELL_MARGIN
    cy0, cy1 = F(gy) * inv_h - CELL_MARGIN, F(gy + 1) * inv_h + CELL_MARGIN
    rej = (cx1 < g[:, 0]) | (g[:, 2] < cx0) | (cy1 < g[:, 1]) | (g[:, 3] < cy0)
    logk = F(np.log((float(anc_w) * float(anc_h)) / (float(img_w) * float(img_h))))
    sp = tw + th + logk
    lg = dm.log(ga.astype(F))
    log_thr = F(np.log(thr))
    rej |= (sp < lg + (log_thr - LOG_MARGIN)) | (sp > lg + (-log_thr + LOG_MARGIN))
    # approximate decode with worst-case corner noise
    sx = F(1) / (F(1) + np.exp(-np.float64(tx))); sy = F(1) / (F(1) + np.exp(-np.float64(ty)))
    fx, fy = (F(sx) + F(gx)) * inv_w, (F(sy) + F(gy)) * inv_h
    fw, fh = F(np.exp(np.float64(tw))) * F(anc_w / img_w), F(np.exp(np.float64(th))) * F(anc_h / img_h)
    noise = rng.uniform(-2e-6, 2e-6, (n, 4)).astype(F)
    fx0, fx1, fy0, fy1 = fx - F(0.5) * fw + noise[:, 0], fx + F(0.5) * fw + noise[:, 1], fy - F(0.5) * fh + noise[:, 2], fy + F(0.5) * fh + noise[:, 3]
    farea = (fx1 - fx0) * (fy1 - fy0)
    iw = np.minimum(fx1, g[:, 2]) - np.maximum(fx0, g[:, 0])
    ih = np.minimum(fy1, g[:, 3]) - np.maximum(fy0, g[:, 1])
    rej |= (iw < F(-1e-5)) | (ih < F(-1e-5))
    fast_ok = (fw <= F(2)) and (fh <= F(2))
    inter = iw * ih
    if fast_ok:
        rej |= (iw >= IOU_FLOOR) & (ih >= IOU_FLOOR) & (inter < (F(thr) - IOU_EPS) * (farea + ga - inter))
    return rej


def test_rejects_never_drop_a_pair_that_reaches_the_threshold():
    rng = np.random.default_rng(20261018)
    image, anchors = 608.0, [(10.0, 13.0), (33.0, 23.0), (116.0, 90.0), (373.0, 326.0)]
    checked = hits = 0
    for case in range(1500):
        W = H = int(rng.choice([19, 38, 76]))
        aw, ah = anchors[int(rng.integers(0, len(anchors)))]
        thr = float(rng.choice([0.5, 0.55, 0.7, 0.9]))
        metric = str(rng.choice(["iou", "diou", "ciou"]))
        gx, gy = int(rng.integers(0, W)), int(rng.integers(0, H))
        tx, ty = F(rng.normal(0, 2)), F(rng.normal(0, 2))
        tw, th = F(np.clip(rng.normal(0, 1.5), -5.9, 4)), F(np.clip(rng.normal(0, 1.5), -5.9, 4))
        # the exact predicted box, as the oracle decodes it (tyu:57-75)
        x = (dm.sigmoid(np.array([tx], F))[0] + F(gx)) / F(W); y = (dm.sigmoid(np.array([ty], F))[0] + F(gy)) / F(H)
        w = dm.exp(np.array([tw], F))[0] * F(aw) / F(image); h = dm.exp(np.array([th], F))[0] * F(ah) / F(image)
        p = np.array([x - w / F(2), y - h / F(2), x + w / F(2), y + h / F(2)], F)
        # ground truth around it: concentric rescalings near the threshold, shifted copies, unrelated boxes
        n = 64
        s = np.sqrt(np.clip(thr * (1 + rng.uniform(-0.04, 0.04, n)), 1e-3, 1.0)) ** rng.choice([-1.0, 1.0], n)
        cx, cy = x + w * rng.normal(0, 0.05, n), y + h * rng.normal(0, 0.05, n)
        gw, gh = w * s * np.exp(rng.normal(0, 0.02, n)), h * s * np.exp(rng.normal(0, 0.02, n))
        far = rng.random(n) < 0.25
        cx[far], cy[far] = rng.random(far.sum()), rng.random(far.sum())
        g = np.stack([cx - gw / 2, cy - gh / 2, cx + gw / 2, cy + gh / 2], 1).astype(F)
        g = g[(g[:, 2] > g[:, 0]) & (g[:, 3] > g[:, 1])]
        exact = oy.get_iou(p.reshape(1, 1, 4), g.reshape(1, -1, 4), metric).reshape(-1)
        rej = _filters_reject(tx, ty, tw, th, gx, gy, W, H, aw, ah, image, image, g, thr, rng)
        reach = (exact >= F(thr)) | np.isnan(exact)
        assert not np.any(rej & reach), "case %d: a pair with metric >= thr was rejected (metric %s, thr %g)" % (case, metric, thr)
        checked += g.shape[0]; hits += int(reach.sum())
    assert checked > 80000 and hits > 8000   # the property was really exercised on both sides of the threshold
