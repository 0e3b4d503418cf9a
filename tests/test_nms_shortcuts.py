"""CPU property tests of the shortcuts in nms.cuh / boxmath.cuh (round 2), restated in NumPy fp32 against the oracle:
  * bm_iou_clearly_below: the division-free reject never fires on a pair whose metric reaches the threshold, for both
    box families and every metric, including pairs planted right at the threshold — and it does dispose of almost all
    overlapping pairs that are below it;
  * the one-warp sweep of the class-agnostic tile: iterating kept' = alive & ~any(by & kept) from kept = alive reaches the
    greedy solution, after at most (longest suppression chain + 1) rounds;
  * the flat pair index space of the per-class buckets: (bucket, i, j) recovered from a pair index by the sqrt formula +
    fix-ups enumerates every pair i < j of every bucket of 2..64 members exactly once, in (j, i) order;
  * nms_bin with estimated bounds (sampled range) stays monotone non-decreasing in the key, so a cut between buckets is a
    cut between keys."""
import numpy as np

from oracle import effdet as oe
from oracle import yolo as oy

F = np.float32


def _clearly_below(iw, ih, s, thr):
    """bm_iou_clearly_below, fp32 operation by operation"""
    thr = F(thr)
    return (iw * ih).astype(F) * (F(1.0) + thr) < (F(0.999) * thr) * s


def _pairs(rng, n, scale):
    """overlapping box pairs (x1,y1,x2,y2), a third of them scaled copies planted around IoU = thr"""
    c = rng.random((n, 2)) * scale
    wh = np.exp(rng.uniform(np.log(0.01), np.log(0.6), (n, 2))) * scale
    a = np.concatenate([c - wh / 2, c + wh / 2], -1)
    d = rng.normal(0, 0.3, (n, 2)) * wh
    wh2 = wh * np.exp(rng.normal(0, 0.5, (n, 2)))
    b = np.concatenate([c + d - wh2 / 2, c + d + wh2 / 2], -1)
    # concentric copies whose IoU is s^2: s around sqrt(thr) for the thresholds used below
    k = n // 3
    s = np.sqrt(rng.choice([0.3, 0.45, 0.5, 0.6], k) * (1.0 + rng.normal(0, 2e-3, k)))[:, None]
    b[:k] = np.concatenate([c[:k] - wh[:k] * s / 2, c[:k] + wh[:k] * s / 2], -1)
    return a.astype(F), b.astype(F)


def test_division_free_reject_is_safe_and_effective():
    rng = np.random.default_rng(11)
    n = 60000
    for scale in (1.0, 512.0):
        a, b = _pairs(rng, n, scale)
        # YOLO family: xyxy, unclamped areas (tf_iou_utils.py:5-65)
        iw = (np.minimum(a[:, 2], b[:, 2]) - np.maximum(a[:, 0], b[:, 0])).astype(F)
        ih = (np.minimum(a[:, 3], b[:, 3]) - np.maximum(a[:, 1], b[:, 1])).astype(F)
        s = ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]) + (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])).astype(F)
        over = (iw > 0) & (ih > 0)
        metrics = {t: oy.get_iou(a, b, t) for t in ("iou", "diou", "ciou")}   # row i of a against row i of b
        for thr in (0.3, 0.45, 0.5, 0.6):
            rej = over & _clearly_below(iw, ih, s, thr)
            for t, m in metrics.items():
                assert not np.any(rej & ~(m < F(thr))), (scale, thr, t)
            below = over & (metrics["iou"] < F(0.99 * thr))
            assert rej[below].mean() > 0.99, (scale, thr)          # what is clearly below is rejected without a division
            near = over & (np.abs(metrics["iou"] / F(thr) - F(1.0)) < F(5e-3))
            assert near.sum() > 500                                  # pairs right at the threshold are exercised
        # EfficientDet family: yxyx, clamped areas (efficientnet/utils/iou.py:26-100)
        ay, by = a[:, [1, 0, 3, 2]], b[:, [1, 0, 3, 2]]
        s2 = (np.maximum(F(0), ay[:, 3] - ay[:, 1]) * np.maximum(F(0), ay[:, 2] - ay[:, 0]) +
              np.maximum(F(0), by[:, 3] - by[:, 1]) * np.maximum(F(0), by[:, 2] - by[:, 0])).astype(F)
        for thr in (0.3, 0.5):
            rej = over & _clearly_below(iw, ih, s2, thr)
            for t in ("iou", "giou", "diou", "ciou"):
                m = oe.get_iou(ay, by, t)
                assert not np.any(rej & ~(m < F(thr))), (scale, thr, t)
            assert rej.sum() > 0.3 * over.sum()


def _greedy(alive, sup):
    kept = np.zeros(64, bool)
    for j in range(64):
        kept[j] = alive[j] and not np.any(kept[:j] & sup[:j, j])
    return kept


def test_fixed_point_sweep_equals_greedy():
    rng = np.random.default_rng(12)
    worst = 0
    for case in range(300):
        dens = rng.choice([0.01, 0.05, 0.2, 0.6])
        sup = np.triu(rng.random((64, 64)) < dens, 1)                 # sup[i, j]: i (better ranked) suppresses j
        if case % 7 == 0:                                               # a long chain: i suppresses i + 1
            sup |= np.triu(np.eye(64, k=1, dtype=bool), 1)
        alive = rng.random(64) < rng.choice([1.0, 0.8, 0.3])
        sup &= alive[:, None] & alive[None, :]                        # rows / columns of dead candidates are zero (phase 2)
        kept = alive.copy()
        rounds = 0
        while True:
            nxt = alive & ~np.any(sup & kept[:, None], axis=0)
            rounds += 1
            if np.array_equal(nxt, kept):
                break
            kept = nxt
            assert rounds <= 65
        worst = max(worst, rounds)
        assert np.array_equal(kept, _greedy(alive, sup)), case
    assert worst > 3                                                    # the chain cases really iterate


def test_flat_pair_index_space():
    rng = np.random.default_rng(13)
    sizes = np.concatenate([rng.integers(0, 65, 200), [64, 2, 1, 0, 64, 70, 300]])   # buckets above 64 hold no pairs
    pv = np.where(sizes <= 64, sizes * (sizes - 1) // 2, 0)
    pb = np.concatenate([[0], np.cumsum(pv)])
    seen = {}
    for e in range(int(pb[-1])):
        b = int(np.searchsorted(pb, e, side="right") - 1)           # sPB[b] <= e < sPB[b + 1]
        assert pv[b] > 0
        r = e - int(pb[b])
        j = int(F(F(1.0) + np.sqrt(F(1.0) + F(8.0) * F(r))) * F(0.5))
        while (j * (j - 1)) >> 1 > r:
            j -= 1
        while ((j + 1) * j) >> 1 <= r:
            j += 1
        i = r - ((j * (j - 1)) >> 1)
        assert 0 <= i < j < sizes[b]
        assert (b, i, j) not in seen
        seen[(b, i, j)] = e
    assert len(seen) == int(pb[-1]) == int(pv.sum())
    # consecutive indices walk (j, i) with i innermost — what the kernel's incremental advance relies on
    items = sorted(seen.items(), key=lambda kv: kv[1])
    for (k0, _), (k1, _) in zip(items, items[1:]):
        if k0[0] == k1[0]:
            assert (k1[2], k1[1]) == ((k0[2], k0[1] + 1) if k0[1] + 1 < k0[2] else (k0[2] + 1, 0))


def _nms_bin(d, dmin, dmax, bins=2048):
    scale = F(bins) / (F(np.uint32(dmax - dmin)) + F(1.0))
    with np.errstate(all="ignore"):
        x = ((d - dmin).astype(np.uint32)).astype(F) * scale          # unsigned wrap below dmin, as on the device
        b = np.where(x >= F(2.0 ** 31), 2 ** 31 - 1, x.astype(np.int64))   # the conversion saturates
    return np.where(d <= dmin, 0, np.minimum(b, bins - 1))


def test_bins_stay_monotone_with_estimated_bounds():
    rng = np.random.default_rng(14)
    for _ in range(50):
        d = np.sort(rng.integers(0, 2 ** 32, 20000, dtype=np.uint64).astype(np.uint32))
        lo, hi = np.sort(rng.integers(0, 2 ** 32, 2, dtype=np.uint64).astype(np.uint32))   # bounds from a "sample": anywhere
        if lo == hi:
            continue
        b = _nms_bin(d, lo, hi)
        assert np.all(np.diff(b) >= 0)
        assert b.min() >= 0 and b.max() <= 2047
