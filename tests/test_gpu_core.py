"""GPU parity: detmath host==device, pairwise metrics and NMS against the NumPy oracle (bit-exact)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F = np.float32


def assert_bits_equal(got, want):
    """Bit-for-bit equality of fp32 arrays; NaNs must coincide but their payload/sign bits are not compared."""
    got = np.asarray(got, F)
    want = np.asarray(want, F)
    assert got.shape == want.shape
    gn, wn = np.isnan(got), np.isnan(want)
    assert np.array_equal(gn, wn), "NaN positions differ"
    ok = np.array_equal(got.view(np.uint32)[~gn], want.view(np.uint32)[~wn])
    if not ok:
        bad = np.flatnonzero((got.view(np.uint32) != want.view(np.uint32)).reshape(-1) & ~gn.reshape(-1))
        raise AssertionError("%d of %d values differ, first at %d: got %r want %r" % (
            bad.size, got.size, bad[0], got.reshape(-1)[bad[0]], want.reshape(-1)[bad[0]]))


def _t(x, cuda, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(cuda)


def test_detmath_device_equals_host_bitwise(lib, cuda):
    import torch
    from oracle import detmath as dm
    rng = np.random.default_rng(11)
    n = 1 << 22
    bits = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    xs_all = bits.view(F)
    xs_all = xs_all[np.isfinite(xs_all)]
    xs_mid = (rng.standard_normal(n) * 8).astype(F)
    cases = [
        (0, dm.exp, np.concatenate([xs_all, xs_mid])),
        (1, dm.sigmoid, np.concatenate([xs_all, xs_mid])),
        (2, dm.log, np.abs(np.concatenate([xs_all, xs_mid]))),
        (3, dm.log1p, np.abs(xs_mid) * F(0.1)),
        (4, dm.atan, np.concatenate([xs_all, xs_mid])),
        (5, lambda v: dm.pow(v, 0.6), rng.random(n, dtype=F)),
        (6, dm.pow15, rng.random(n, dtype=F)),
    ]
    for op, host_fn, x in cases:
        x = np.ascontiguousarray(x, dtype=F)
        dx = _t(x, cuda)
        out = torch.empty_like(dx)
        st = lib.b200_detmath_eval(op, dx.data_ptr(), 0, out.data_ptr(), x.size, torch.cuda.current_stream().cuda_stream)
        assert st == 0, lib.b200_last_error()
        got = out.cpu().numpy()
        want = host_fn(x)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "op %d differs in %d places" % (
            op, int((got.view(np.uint32) != want.view(np.uint32)).sum()))
    x = xs_mid
    z = rng.random(n, dtype=F)
    dx, dz = _t(x, cuda), _t(z, cuda)
    out = torch.empty_like(dx)
    assert lib.b200_detmath_eval(7, dx.data_ptr(), dz.data_ptr(), out.data_ptr(), n, torch.cuda.current_stream().cuda_stream) == 0
    assert np.array_equal(out.cpu().numpy().view(np.uint32), dm.bce_logits(z, x).view(np.uint32))


def _rand_boxes(rng, n, scale=1.0, yx=False, cluster=0):
    c = rng.random((n, 2)) * scale
    if cluster:
        centres = rng.random((cluster, 2)) * scale
        c = centres[rng.integers(0, cluster, n)] + rng.normal(0, 0.01 * scale, (n, 2))
    wh = np.exp(rng.uniform(np.log(0.02), np.log(0.4), (n, 2))) * scale
    b = np.concatenate([c - wh / 2, c + wh / 2], -1).astype(F)
    return b[:, [1, 0, 3, 2]].copy() if yx else b


@pytest.mark.parametrize("iou_type", ["iou", "diou", "ciou"])
def test_get_iou_yolo_bitwise(lib, cuda, iou_type):
    from oracle import yolo as oy
    from tfmv_b200.ai_models.utils.tf_iou_utils import GetIOU
    rng = np.random.default_rng(21)
    b1 = _rand_boxes(rng, 6 * 5 * 3).reshape(6, 5, 3, 1, 4)
    b2 = _rand_boxes(rng, 37)[None]
    b2[0, 3] = b1[0, 0, 0, 0]  # identical pair
    b2[0, 4] = [0.2, 0.2, 0.2, 0.2]  # zero-size
    want = oy.get_iou(b1, b2, iou_type)
    got = GetIOU(_t(b1, cuda), _t(b2, cuda), iou_type).cpu().numpy()
    assert got.shape == want.shape == (6, 5, 3, 37)
    assert_bits_equal(got, want)
    # elementwise broadcast form used inside the NMS loops: (1,4) x (n,4)
    w2 = oy.get_iou(b1.reshape(-1, 4)[0:1], b2[0], iou_type)
    g2 = GetIOU(_t(b1.reshape(-1, 4)[0:1], cuda), _t(b2[0], cuda), iou_type).cpu().numpy()
    assert_bits_equal(g2, w2)
    with pytest.raises(AssertionError):
        GetIOU(_t(b1, cuda), _t(b2, cuda), "giou")


@pytest.mark.parametrize("iou_type", ["iou", "giou", "diou", "ciou"])
def test_get_iou_effdet_bitwise(lib, cuda, iou_type):
    from oracle import effdet as oe
    from tfmv_b200.ai_models.efficientnet.utils.iou import get_iou
    rng = np.random.default_rng(22)
    b1 = _rand_boxes(rng, 4 * 4 * 9, 512, yx=True).reshape(4, 4, 9, 1, 4)
    b2 = _rand_boxes(rng, 23, 512, yx=True)
    b2[5] = b1[1, 1, 1, 0]
    b2[6] = [5, 5, 5, 5]
    b2[7] = [9, 9, 3, 3]  # inverted -> clamped to zero area
    want = oe.get_iou(b1, b2, iou_type)
    got = get_iou(_t(b1, cuda), _t(b2, cuda), iou_type).cpu().numpy()
    assert_bits_equal(got, want)


NMS_CASES = [
    # n, cluster, classes, max_out, thr
    (1, 0, 3, 10, 0.5), (2, 0, 1, 10, 0.5), (63, 4, 2, 500, 0.5), (64, 0, 80, 500, 0.5), (65, 8, 3, 500, 0.3),
    (700, 40, 5, 500, 0.5), (1024, 30, 4, 500, 0.5), (1025, 100, 80, 500, 0.45), (3000, 50, 3, 500, 0.5),
    (5000, 0, 80, 500, 0.5), (2500, 20, 1, 2000, 0.6), (1500, 10, 2, 7, 0.5),
]


@pytest.mark.parametrize("iou_type", ["iou", "diou", "ciou"])
@pytest.mark.parametrize("case", NMS_CASES)
def test_nms_yolo_family_matches_oracle(lib, cuda, iou_type, case):
    from oracle import yolo as oy
    from tfmv_b200.ai_models.utils.tf_iou_utils import GetIOUNMS, GetIOUNMSByClasses
    n, cluster, ncls, max_out, thr = case
    rng = np.random.default_rng(1000 + n)
    boxes = _rand_boxes(rng, n, cluster=cluster)
    scores = rng.random(n, dtype=F)
    scores[rng.integers(0, n, n // 3)] = F(0.75)  # many exact ties
    if n > 10:
        boxes[5] = boxes[2]
        boxes[7] = [0.5, 0.5, 0.5, 0.5]
    classes = rng.integers(0, ncls, n).astype(np.int32)
    want = oy.get_iou_nms_by_classes(boxes, scores, classes, max_out, thr, iou_type)
    got = GetIOUNMSByClasses(_t(boxes, cuda), _t(scores, cuda), _t(classes, cuda), max_out, thr, iou_type).cpu().numpy()
    assert got.dtype == np.int32 and got.tolist() == want.tolist()
    want = oy.get_iou_nms(boxes, scores, max_out, thr, iou_type)
    got = GetIOUNMS(_t(boxes, cuda), _t(scores, cuda), max_out, thr, iou_type).cpu().numpy()
    assert got.tolist() == want.tolist()


@pytest.mark.parametrize("iou_type", ["iou", "ciou"])
@pytest.mark.parametrize("n,cluster,max_out", [(900, 60, 500), (2600, 90, 700), (6000, 200, 500)])
def test_nms_by_class_mixed_bucket_sizes_and_shared_buckets(lib, cuda, iou_type, n, cluster, max_out):
    """Per-class NMS resolves class buckets (class id & 255) two ways: buckets of up to 64 members from pairwise suppression
    words, larger ones by a warp loop.  Skewed class frequencies put both kinds — and single-member buckets — into one chunk;
    class ids 256 apart share a bucket and must not interact."""
    from oracle import yolo as oy
    from tfmv_b200.ai_models.utils.tf_iou_utils import GetIOUNMSByClasses
    rng = np.random.default_rng(77 + n)
    boxes = _rand_boxes(rng, n, cluster=cluster)
    scores = rng.random(n, dtype=F)
    scores[rng.integers(0, n, n // 4)] = F(0.5)   # exact ties
    ids = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 17, 256, 257, 258 + 256, 300, 1 + 512, 40, 41, 42, 43], dtype=np.int32)
    prob = np.array([30, 20, 10, 5, 3, 2, 2, 2, 2, 2, 1, 6, 5, 2, 1, 4, 1, 1, 0.5, 0.5])
    classes = ids[rng.choice(len(ids), n, p=prob / prob.sum())].astype(np.int32)
    want = oy.get_iou_nms_by_classes(boxes, scores, classes, max_out, 0.45, iou_type)
    got = GetIOUNMSByClasses(_t(boxes, cuda), _t(scores, cuda), _t(classes, cuda), max_out, 0.45, iou_type).cpu().numpy()
    assert got.tolist() == want.tolist()
    assert 0 < len(want) <= max_out


@pytest.mark.parametrize("iou_type", ["iou", "giou", "diou", "ciou"])
@pytest.mark.parametrize("n,cluster,max_out,sthr", [(1, 0, 200, None), (300, 30, 200, 0.2), (4000, 300, 200, 1e-4),
                                                      (20000, 0, 200, 1e-4), (2000, 3, 200, -5.0)])
def test_nms_effdet_matches_oracle(lib, cuda, iou_type, n, cluster, max_out, sthr):
    from oracle import effdet as oe
    from tfmv_b200.ai_models.efficientnet.utils.nms import get_nms
    rng = np.random.default_rng(2000 + n)
    boxes = _rand_boxes(rng, n, 512, yx=True, cluster=cluster)
    scores = rng.standard_normal(n).astype(F)
    scores[rng.integers(0, n, n // 4)] = F(1.25)
    kw = {} if sthr is None else {"score_threshold": sthr}
    want = oe.get_nms(boxes, scores, max_out, 0.5, iou_type=iou_type, **({"score_threshold": sthr} if sthr is not None else {}))
    got = get_nms(_t(boxes, cuda), _t(scores, cuda), max_out, 0.5, iou_type=iou_type, **kw).cpu().numpy()
    assert got.tolist() == want.tolist()


def test_nms_batched_segments_and_order_ids(lib, cuda):
    """Segments are independent; order_id decides ties regardless of storage order (used by the fused decode)."""
    import torch
    from oracle import yolo as oy
    rng = np.random.default_rng(77)
    sizes = [0, 5, 1300, 1, 257]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    tot = int(off[-1])
    boxes = _rand_boxes(rng, tot, cluster=25)
    scores = np.round(rng.random(tot, dtype=F), 2)  # coarse -> lots of ties
    classes = rng.integers(0, 4, tot).astype(np.int32)
    perm_all, oid = [], np.empty(tot, np.uint32)
    for s in range(len(sizes)):
        p = rng.permutation(sizes[s])
        perm_all.append(p + off[s])
        oid[off[s]:off[s + 1]] = p  # stored slot j holds original element p[j]
    perm_all = np.concatenate(perm_all).astype(np.int64) if tot else np.zeros(0, np.int64)
    sb, ss, sc = boxes[perm_all], scores[perm_all], classes[perm_all]
    max_out = 300
    out_idx = torch.full((len(sizes), max_out), -1, dtype=torch.int32, device=cuda)
    out_cnt = torch.zeros((len(sizes),), dtype=torch.int32, device=cuda)
    d = lambda a, dt=None: _t(a, cuda, dt)
    dsb, dss, dsc, doid, doff = d(sb), d(ss), d(sc), d(oid.view(np.int32)), d(off)
    st = lib.b200_nms(dsb.data_ptr(), dss.data_ptr(), dsc.data_ptr(), doid.data_ptr(), doff.data_ptr(), len(sizes),
                      1, 1, ctypes.c_float(0.5), 0, ctypes.c_float(0.0), max_out, out_idx.data_ptr(), out_cnt.data_ptr(),
                      torch.cuda.current_stream().cuda_stream)
    assert st == 0, lib.b200_last_error()
    cnt = out_cnt.cpu().numpy()
    idx = out_idx.cpu().numpy()
    for s in range(len(sizes)):
        a, b = off[s], off[s + 1]
        want = oy.get_iou_nms_by_classes(boxes[a:b], scores[a:b], classes[a:b], max_out, 0.5, "diou")
        got_orig = oid[a:b][idx[s, :cnt[s]]] if cnt[s] else np.zeros(0, np.int64)
        assert got_orig.tolist() == want.tolist(), "segment %d" % s
