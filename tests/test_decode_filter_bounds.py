"""CPU property tests of the shortcuts in yolo_decode_filter_kernel (csrc/yolo_decode.cu), restated in NumPy against
the oracle's deterministic sigmoid / exp:
  * conf threshold decided in logit space outside [conf_lo, conf_hi];
  * class maximum: a logit more than 0.01 below the maximum (maximum in (-80, 8)) has a strictly smaller sigmoid;
  * box validity (x2 > x1 and y2 > y1) is certain for tw, th in [tmin, 80] with non-NaN tx, ty."""
import numpy as np

from oracle import detmath as dm

F = np.float32


def _conf_bounds(thr):
    """host code of b200_yolo_decode_nms"""
    if not (F(thr) >= F(1e-3) and F(thr) <= F(1.0) - F(1e-3)):
        return -np.inf, np.inf
    t = np.log(float(F(thr)) / (1.0 - float(F(thr))))
    d = 3e-6 / (float(F(thr)) * (1.0 - float(F(thr)))) + 1e-6 * abs(t)
    return F(t - d), F(t + d)


def test_conf_threshold_in_logit_space():
    rng = np.random.default_rng(1)
    for thr in [0.5, 0.3, 0.7, 0.05, 0.95, 0.001, 0.999, 0.25, 0.9]:
        lo, hi = _conf_bounds(thr)
        t = np.log(thr / (1 - thr))
        x = np.concatenate([t + rng.normal(0, 1e-5, 20000), t + rng.normal(0, 1e-3, 20000), rng.normal(0, 4, 20000),
                            np.nextafter(F(hi), F(np.inf), dtype=F).repeat(4), np.nextafter(F(lo), F(-np.inf), dtype=F).repeat(4)]).astype(F)
        s = dm.sigmoid(x)
        assert np.all(s[x > hi] > F(thr)), thr
        assert not np.any(s[x < lo] > F(thr)), thr
        assert np.any((x >= lo) & (x <= hi))  # the exact band is exercised too


def test_class_guard_band():
    rng = np.random.default_rng(2)
    m = np.concatenate([rng.uniform(-79.9, 7.99, 200000), rng.uniform(6.0, 7.999, 50000), rng.uniform(-79.99, -60, 50000)]).astype(F)
    gap = np.concatenate([np.full(100000, 0.01), rng.uniform(0.01, 0.02, 100000), rng.uniform(0.01, 5.0, 100000)]).astype(F)
    x = (m - gap).astype(F)
    keep = x < m - F(0.01)           # the kernel's test: runner-up r2 < lo = m1 - 0.01f (fp32 arithmetic)
    sm, sx = dm.sigmoid(m), dm.sigmoid(x)
    assert np.all(sx[keep] < sm[keep])
    assert keep.sum() > 100000


def test_validity_safe_range():
    rng = np.random.default_rng(3)
    n = 200000
    image = F(416)
    anc = np.array([10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326], F).reshape(9, 2)
    a = rng.integers(0, 9, n)
    aw, ah = anc[a, 0] / image, anc[a, 1] / image
    tmin_w = (np.log(4e-7 / aw.astype(np.float64)) + 1e-3).astype(F)
    tmin_h = (np.log(4e-7 / ah.astype(np.float64)) + 1e-3).astype(F)
    # logits at and inside the edges of the safe range
    u = rng.random(n)
    tw = np.where(u < 0.3, tmin_w, np.where(u < 0.6, F(80.0), rng.uniform(-20, 80, n))).astype(F)
    th = np.where(rng.random(n) < 0.3, tmin_h, rng.uniform(-20, 80, n)).astype(F)
    tx, ty = (rng.standard_normal(n) * 20).astype(F), (rng.standard_normal(n) * 20).astype(F)
    W = rng.choice([13, 26, 52, 76], n)
    gx, gy = (rng.random(n) * W).astype(int), (rng.random(n) * W).astype(int)
    safe = (tw >= tmin_w) & (tw <= F(80)) & (th >= tmin_h) & (th <= F(80))
    with np.errstate(all="ignore"):
        x = (dm.sigmoid(tx) + gx.astype(F)) / W.astype(F)
        y = (dm.sigmoid(ty) + gy.astype(F)) / W.astype(F)
        w = dm.exp(tw) * aw
        h = dm.exp(th) * ah
        w = np.where(np.isinf(w), F(0), w); h = np.where(np.isinf(h), F(0), h)
        valid = ((x + w / F(2)) > (x - w / F(2))) & ((y + h / F(2)) > (y - h / F(2)))
    assert np.all(valid[safe])
    assert safe.sum() > 50000 and (~valid).sum() > 100   # both sides exercised
