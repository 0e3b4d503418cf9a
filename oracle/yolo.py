"""NumPy oracle for the YOLOv3/v4 part of the path — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Each function restates one reference function in fp32, one NumPy op per TensorFlow op and in the
same order, so that + - * / max min steps round exactly as TF's per-op kernels do.  Paths cited are
relative to /root/reference/AIServer/ai_api/ai_models/.
"""
import numpy as np

from . import detmath as dm

F = np.float32
COCO_ANCHORS_FLAT = np.array(
    [10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326], dtype=np.float32
).reshape(9, 2)  # yolo_v3/unit_test/loss_test.py:154-156


def load_anchors_order(flat18):
    """utils/load_object_detection_data.py:58-67 — reshape (3,-1,2) then reorder [2,1,0]: layer 0 = largest."""
    a = np.asarray(flat18, dtype=np.int64).reshape(3, -1, 2)
    return a[[2, 1, 0]]


# --------------------------------------------------------------------------------------------
def grid_meshgrid(h, w):
    """utils/tf_yolo_utils.py:28-33 — (H,W,1,2) with [...,0]=x (fastest), [...,1]=y."""
    gx, gy = np.meshgrid(np.arange(w, dtype=F), np.arange(h, dtype=F))
    return np.concatenate([gx.reshape(h, w, 1, 1), gy.reshape(h, w, 1, 1)], axis=-1)


def grid_tile(h, w):
    """losses/yolo_loss.py:62-68 — the keras-yolo3 tile/arange construction."""
    gy = np.tile(np.arange(h).reshape(-1, 1, 1, 1), [1, w, 1, 1])
    gx = np.tile(np.arange(w).reshape(1, -1, 1, 1), [h, 1, 1, 1])
    return np.concatenate([gx, gy], axis=-1).astype(F)


# --------------------------------------------------------------------------------------------
def get_iou(b1, b2, iou_type="iou"):
    """utils/tf_iou_utils.py:5-65 GetIOU. b1 (..., n1, 1, 4), b2 (1, n2, 4), corners x1,y1,x2,y2."""
    assert iou_type in ["iou", "diou", "ciou"]
    b1 = np.asarray(b1, dtype=F)
    b2 = np.asarray(b2, dtype=F)
    with np.errstate(all="ignore"):
        imin = np.maximum(b1[..., 0:2], b2[..., 0:2])
        imax = np.minimum(b1[..., 2:4], b2[..., 2:4])
        iwh = np.maximum(imax - imin, F(0.0))
        inter = iwh[..., 0] * iwh[..., 1]
        wh1 = b1[..., 2:4] - b1[..., 0:2]
        wh2 = b2[..., 2:4] - b2[..., 0:2]
        a1 = wh1[..., 0] * wh1[..., 1]
        a2 = wh2[..., 0] * wh2[..., 1]
        iou = inter / ((a1 + a2) - inter)
        if iou_type == "iou":
            return iou
        umin = np.minimum(b1[..., 0:2], b2[..., 0:2])
        umax = np.maximum(b1[..., 2:4], b2[..., 2:4])
        uwh = umax - umin
        c = np.square(uwh[..., 0]) + np.square(uwh[..., 1])
        c1 = (b1[..., 2:4] + b1[..., 0:2]) / F(2)
        c2 = (b2[..., 2:4] + b2[..., 0:2]) / F(2)
        dd = np.square(c1 - c2)
        u = dd[..., 0] + dd[..., 1]
        d = u / c
        diou = iou - dm.pow(d, 0.6)
        diou = np.where(c == F(0.0), iou, diou)
        if iou_type == "diou":
            return diou
        coef = F(4) / np.square(F(np.pi))
        at = dm.atan(wh1[..., 0] / wh1[..., 1]) - dm.atan(wh2[..., 0] / wh2[..., 1])
        v = coef * np.square(at)
        alpha = v / (((F(1) - iou) + v) + F(1e-8))
        ciou = iou - (d + alpha * v)
        return np.where(c == F(0.0), iou, ciou)


def _argsort_desc(scores):
    """tf.argsort(direction='DESCENDING') = top_k(k=n): equal scores keep the lower index first."""
    return np.argsort(-np.asarray(scores, dtype=F), kind="stable").astype(np.int32)


def get_iou_nms(boxes, scores, max_output_size, iou_threshold=0.5, iou_type="iou"):
    """utils/tf_iou_utils.py:67-108 GetIOUNMS — class-agnostic greedy; survivors need metric < thr (NaN drops)."""
    boxes = np.asarray(boxes, F).reshape(-1, 4)
    order = _argsort_desc(scores)
    cur = boxes[order]
    out = []
    thr = F(iou_threshold)
    while len(out) < max_output_size and cur.shape[0] > 0:
        out.append(int(order[0]))
        if cur.shape[0] == 1:
            break
        m = get_iou(cur[0:1], cur[1:], iou_type)
        with np.errstate(all="ignore"):
            keep = m < thr
        cur = cur[1:][keep]
        order = order[1:][keep]
    return np.asarray(out, dtype=np.int32)


def get_iou_nms_by_classes(boxes, scores, classes, max_output_size, iou_threshold=0.5, iou_type="iou"):
    """utils/tf_iou_utils.py:110-157 GetIOUNMSByClasses — survivors: not(metric >= thr and same class) (NaN stays)."""
    boxes = np.asarray(boxes, F).reshape(-1, 4)
    classes = np.asarray(classes)
    order = _argsort_desc(scores)
    cur = boxes[order]
    cls = classes[order]
    out = []
    thr = F(iou_threshold)
    while len(out) < max_output_size and cur.shape[0] > 0:
        out.append(int(order[0]))
        if cur.shape[0] == 1:
            break
        m = get_iou(cur[0:1], cur[1:], iou_type)
        with np.errstate(all="ignore"):
            keep = np.logical_not(np.logical_and(m >= thr, cls[1:] == cls[0]))
        cur = cur[1:][keep]
        order = order[1:][keep]
        cls = cls[1:][keep]
    return np.asarray(out, dtype=np.int32)


# --------------------------------------------------------------------------------------------
def get_boxes(y, anchors_wh, classes_num, return_mask=False):
    """utils/tf_yolo_utils.py:129-167 GetBoxes. y (B,H,W,A,5+C); anchors_wh (A,2) already / image_wh."""
    y = np.asarray(y, F)
    anchors_wh = np.asarray(anchors_wh, F)
    h, w = y.shape[1], y.shape[2]
    conf = dm.sigmoid(y[..., 4:5])
    classes = dm.sigmoid(y[..., 5 : 5 + classes_num])
    gx, gy = np.meshgrid(np.arange(w), np.arange(h))
    grid = np.concatenate([gx.reshape(h, w, 1, 1), gy.reshape(h, w, 1, 1)], axis=-1).astype(F)
    grid_wh = np.array([w, h], dtype=F)
    xy = (dm.sigmoid(y[..., 0:2]) + grid) / grid_wh
    with np.errstate(all="ignore"):
        wh = dm.exp(y[..., 2:4]) * anchors_wh
    wh = np.where(np.isinf(wh), F(0.0), wh)
    half = wh / F(2)
    boxes = np.concatenate([xy - half, xy + half], axis=-1)
    mask = np.logical_and(boxes[..., 2] > boxes[..., 0], boxes[..., 3] > boxes[..., 1])
    if return_mask:
        return boxes[mask], conf[mask], classes[mask], mask
    return boxes[mask], conf[mask], classes[mask]


def get_nms_boxes_ex(y1, y2, y3, anchors_wh, image_wh, classes_num, confidence_thresh=0.5, scores_thresh=0.3,
                     iou_thresh=0.5, iou_type="iou", max_output_size=500):
    """utils/tf_yolo_utils.py:169-269 GetNMSBoxes, plus the intermediate candidate list for parity tests.

    Returns dict with the reference's five outputs (boxes, classes_id, scores, classes, confidence),
    `selected` (indices into the compacted candidate list, emit order) and `cand_anchor` (flat anchor
    index level-major / h / w / a of every candidate) for checking the GPU's compaction order.
    The reference flattens the batch (tyu:163-166); this is defined behaviour for B == 1 only.
    """
    image_wh_f = np.asarray(image_wh, F)
    anchors_wh_f = np.asarray(anchors_wh, F)
    a_num = anchors_wh_f.shape[1]
    ys = []
    for y in (y1, y2, y3):
        y = np.asarray(y, F)
        ys.append(y.reshape(y.shape[0], y.shape[1], y.shape[2], a_num, -1))
    cb, cid, cs, cc, cf, ca = [], [], [], [], [], []
    base = 0
    for l, y in enumerate(ys):
        boxes, conf, classes, vmask = get_boxes(y, anchors_wh_f[l] / image_wh_f, classes_num, return_mask=True)
        flat_idx = np.flatnonzero(vmask.reshape(-1)) + base
        base += vmask.size
        cmax = np.max(classes, axis=-1, keepdims=True) if classes.shape[0] else np.zeros((0, 1), F)
        m = np.logical_and(conf > F(confidence_thresh), cmax > F(scores_thresh))[..., 0]
        boxes, classes, conf, flat_idx = boxes[m], classes[m], conf[m], flat_idx[m]
        cb.append(boxes.reshape(-1, 4))
        cs.append(np.max(classes, axis=-1).reshape(-1) if classes.shape[0] else np.zeros((0,), F))
        cid.append(np.argmax(classes, axis=-1).astype(np.int32).reshape(-1) if classes.shape[0] else np.zeros((0,), np.int32))
        cc.append(classes)
        cf.append(conf)
        ca.append(flat_idx.astype(np.int64))
    cb, cid, cs = np.concatenate(cb, 0), np.concatenate(cid, 0), np.concatenate(cs, 0)
    cc, cf, ca = np.concatenate(cc, 0), np.concatenate(cf, 0), np.concatenate(ca, 0)
    sel = get_iou_nms_by_classes(cb, cs, cid, max_output_size, iou_threshold=iou_thresh, iou_type=iou_type)
    return dict(boxes=cb[sel], classes_id=cid[sel], scores=cs[sel], classes=cc[sel], confidence=cf[sel],
                selected=sel, cand_anchor=ca, cand_boxes=cb, cand_scores=cs, cand_classes_id=cid)


def get_nms_boxes(*args, **kwargs):
    r = get_nms_boxes_ex(*args, **kwargs)
    return r["boxes"], r["classes_id"], r["scores"], r["classes"], r["confidence"]


# --------------------------------------------------------------------------------------------
def _sum32(x):
    """tf.reduce_sum of an fp32 tensor; TF's summation order is unspecified, so accumulate in fp64."""
    return F(np.sum(np.asarray(x, dtype=np.float64)))


def get_loss(y_true, y_pred, image_wh, anchors_wh, iou_thresh=0.5, iou_type="iou", return_parts=False,
             variant="tf_yolo_utils", return_ignore=False):
    """utils/tf_yolo_utils.py:6-127 GetLoss.

    variant="unit_test_copy" reproduces the local copy in yolo_v3/unit_test/loss_test.py:18-150
    (no +1e-8 in the log, raw_true_xy not multiplied by obj) so the reference's own relation
    GetLoss-copy == Yolov4Loss can be re-checked on this oracle.
    """
    image_wh_f = np.asarray(image_wh, F)
    anchors_wh_f = np.asarray(anchors_wh, F)
    bsz = np.asarray(y_true[0]).shape[0]
    bf = F(bsz)
    loss = F(0.0)
    parts = []
    ignores = []
    for l in range(3):
        yt = np.asarray(y_true[l], F)
        yp = np.asarray(y_pred[l], F).reshape(yt.shape)
        h, w = yt.shape[1], yt.shape[2]
        grid = grid_meshgrid(h, w)
        grid_wh = np.array([w, h], dtype=F)
        obj = yt[..., 4:5]
        t_cls = yt[..., 5:]
        t_xy = yt[..., 0:2]
        raw_xy = t_xy * grid_wh - grid
        if variant == "tf_yolo_utils":
            raw_xy = obj * raw_xy
        t_wh = yt[..., 2:4]
        with np.errstate(all="ignore"):
            if variant == "tf_yolo_utils":
                raw_wh = dm.log((t_wh * image_wh_f[::-1] + F(1e-8)) / anchors_wh_f[l])
            else:
                raw_wh = dm.log(t_wh * image_wh_f[::-1] / anchors_wh_f[l])
        raw_wh = np.where(obj.astype(bool), raw_wh, F(0.0))
        p_obj = yp[..., 4:5]
        p_cls = yp[..., 5:]
        p_xy_raw = yp[..., 0:2]
        p_xy = (dm.sigmoid(p_xy_raw) + grid) / grid_wh
        p_wh_raw = yp[..., 2:4]
        with np.errstate(all="ignore"):
            p_wh = dm.exp(p_wh_raw) * anchors_wh_f[l] / image_wh_f[::-1]
        t_half = t_wh / F(2)
        t_boxes = np.concatenate([t_xy - t_half, t_xy + t_half], axis=-1)
        p_half = p_wh / F(2)
        p_boxes = np.concatenate([p_xy - p_half, p_xy + p_half], axis=-1)
        ignore = np.empty(yt.shape[:4], dtype=F)
        for b in range(bsz):
            gt = t_boxes[b][obj[b, ..., 0] != 0]  # tf.boolean_mask with a float mask == nonzero
            m = get_iou(p_boxes[b][..., None, :], gt[None, ...], iou_type=iou_type)
            if gt.shape[0] == 0:
                best = np.full(m.shape[:-1], np.finfo(F).min, dtype=F)  # reduce_max over empty axis
            else:
                best = np.max(m, axis=-1)
            with np.errstate(all="ignore"):
                ignore[b] = (best < F(iou_thresh)).astype(F)
        ignores.append(ignore.reshape(bsz, -1).astype(np.uint8))
        ignore = ignore[..., None]
        scale = F(2) - t_wh[..., 0:1] * t_wh[..., 1:2]
        xy_bc = dm.bce_logits(raw_xy, p_xy_raw)
        xy_loss = obj * scale * xy_bc
        wh_loss = obj * scale * F(0.5) * np.square(raw_wh - p_wh_raw)
        obj_bc = dm.bce_logits(obj, p_obj)
        obj_loss = obj * obj_bc + (F(1) - obj) * obj_bc * ignore
        cls_bc = dm.bce_logits(t_cls, p_cls)
        cls_loss = obj * cls_bc
        s = [_sum32(xy_loss) / bf, _sum32(wh_loss) / bf, _sum32(obj_loss) / bf, _sum32(cls_loss) / bf]
        parts.append([F(v) for v in s])
        loss = F(loss + F(F(F(s[0] + s[1]) + s[2]) + s[3]))
    if return_ignore:
        return loss, np.asarray(parts, dtype=F), np.concatenate(ignores, axis=1)
    if return_parts:
        return loss, np.asarray(parts, dtype=F)
    return loss


def yolov4_loss(anchors9, classes_num, y_true, y_pred, ignore_thresh=0.5):
    """losses/yolo_loss.py:85-159 Yolov4Loss.call (anchors flat ascending (9,2), anchor_mask [[6,7,8],[3,4,5],[0,1,2]])."""
    anchors9 = np.asarray(anchors9)
    mask = [[6, 7, 8], [3, 4, 5], [0, 1, 2]]
    h0, w0 = np.asarray(y_pred[0]).shape[1:3]
    input_shape = np.array([h0 * 32, w0 * 32], dtype=F)  # (H, W)
    bsz = np.asarray(y_pred[0]).shape[0]
    mf = F(bsz)
    loss = F(0.0)
    for l in range(3):
        yt = np.asarray(y_true[l], F)
        h, w = yt.shape[1], yt.shape[2]
        grid_hw = np.array([h, w], dtype=F)
        anc = anchors9[mask[l]].astype(F).reshape(1, 1, 1, 3, 2)
        grid = grid_tile(h, w)
        feats = np.asarray(y_pred[l], F).reshape(-1, h, w, 3, classes_num + 5)
        pred_xy = (dm.sigmoid(feats[..., :2]) + grid) / grid_hw[::-1]
        with np.errstate(all="ignore"):
            pred_wh = dm.exp(feats[..., 2:4]) * anc / input_shape[::-1]
        pred_box = np.concatenate([pred_xy, pred_wh], axis=-1)
        obj = yt[..., 4:5]
        t_cls = yt[..., 5:]
        raw_xy = yt[..., :2] * grid_hw[::-1] - grid
        with np.errstate(all="ignore"):
            raw_wh = dm.log(yt[..., 2:4] * input_shape[::-1] / anchors9[mask[l]].astype(F))
        raw_wh = np.where(obj.astype(bool), raw_wh, F(0.0))
        scale = F(2) - yt[..., 2:3] * yt[..., 3:4]
        ignore = np.empty(yt.shape[:4], dtype=F)
        for b in range(bsz):
            tb = yt[b, ..., 0:4][obj[b, ..., 0].astype(bool)]
            # BoxIou on xywh boxes, losses/yolo_loss.py:13-51
            b1 = pred_box[b][..., None, :]
            b1h = b1[..., 2:4] / F(2.0)
            b1min, b1max = b1[..., :2] - b1h, b1[..., :2] + b1h
            b2 = tb[None, ...]
            b2h = b2[..., 2:4] / F(2.0)
            b2min, b2max = b2[..., :2] - b2h, b2[..., :2] + b2h
            with np.errstate(all="ignore"):
                iwh = np.maximum(np.minimum(b1max, b2max) - np.maximum(b1min, b2min), F(0.0))
                inter = iwh[..., 0] * iwh[..., 1]
                a1 = b1[..., 2] * b1[..., 3]
                a2 = b2[..., 2] * b2[..., 3]
                iou = inter / ((a1 + a2) - inter)
                best = np.max(iou, axis=-1) if tb.shape[0] else np.full(iou.shape[:-1], np.finfo(F).min, F)
                ignore[b] = (best < F(ignore_thresh)).astype(F)
        ignore = ignore[..., None]
        xy_loss = obj * scale * dm.bce_logits(raw_xy, feats[..., 0:2])
        wh_loss = obj * scale * F(0.5) * np.square(raw_wh - feats[..., 2:4])
        cbc = dm.bce_logits(obj, feats[..., 4:5])
        conf_loss = obj * cbc + (F(1) - obj) * cbc * ignore
        cls_loss = obj * dm.bce_logits(t_cls, feats[..., 5:])
        s = [_sum32(xy_loss) / mf, _sum32(wh_loss) / mf, _sum32(conf_loss) / mf, _sum32(cls_loss) / mf]
        loss = F(loss + F(F(F(s[0] + s[1]) + s[2]) + s[3]))
    return loss


# --------------------------------------------------------------------------------------------
def get_targets(boxes, classes, anchors_wh, image_wh, classes_num, layers_hw=None):
    """datasets/coco_dataset.py:185-285 DataGenerator.GetTargets for one image.

    boxes (n,4) pixel corners x1,y1,x2,y2; classes (n,) int; anchors_wh (3,A,2) pixels, layer 0 = stride 32.
    Returns 3 dense targets (H,W,A,5+C).  Reproduces: float floor-div centre (cds:193), normalised-wh vs
    pixel-anchor IoU (cds:200-219), layer = idx // layers_num and anchor = idx % layers_num (cds:237,241),
    scatter_nd duplicate sums (cds:265-276) and zeroing of every record whose obj > 1 (cds:279-284).
    """
    boxes = np.asarray(boxes, F).reshape(-1, 4)
    classes = np.asarray(classes).reshape(-1)
    anchors_wh = np.asarray(anchors_wh)
    image_wh_f = np.asarray(image_wh, F)
    layers_num = anchors_wh.shape[0]
    a_num = anchors_wh.shape[1]
    if layers_hw is None:
        layers_hw = [[int(image_wh[1]) // s, int(image_wh[0]) // s] for s in (32, 16, 8)]  # cds:54
    xy = np.floor_divide(boxes[:, 2:4] + boxes[:, 0:2], F(2))
    wh = boxes[:, 2:4] - boxes[:, 0:2]
    xy = xy / image_wh_f
    wh = wh / image_wh_f
    bmax = wh / F(2.0)
    bmm = np.concatenate([-bmax, bmax], axis=-1)[:, None, :]
    amax = anchors_wh.reshape(-1, 2).astype(F) / F(2.0)
    abox = np.concatenate([-amax, amax], axis=-1)[None, ...]
    iou = get_iou(bmm, abox, "iou")
    aidx = np.argmax(iou, axis=-1).astype(np.int32) if boxes.shape[0] else np.zeros((0,), np.int32)
    targets = [np.zeros((layers_hw[l][0], layers_hw[l][1], a_num, 5 + classes_num), dtype=F) for l in range(3)]
    for i in range(boxes.shape[0]):
        layer = int(aidx[i]) // layers_num
        anchor = int(aidx[i]) % layers_num
        yx = np.floor(xy[i][::-1] * np.asarray(layers_hw[layer], dtype=F)).astype(np.int32)
        onehot = np.zeros((classes_num,), dtype=F)
        if 0 <= int(classes[i]) < classes_num:
            onehot[int(classes[i])] = F(1.0)
        upd = np.concatenate([xy[i], wh[i], np.array([1.0], F), onehot]).astype(F)
        targets[layer][yx[0], yx[1], anchor] += upd  # scatter_nd sums duplicates
    out = []
    for t in targets:
        keep = (t[..., 4:5] <= F(1)).astype(F)
        out.append(t * keep)
    return tuple(out)


def get_ground_truth(y):
    """yolo_v4/model.py:380-395 GetGroudTruth: dense targets -> (n,5) [x1,y1,x2,y2,class]."""
    y = np.asarray(y, F)
    conf = y[..., 4]
    half = y[..., 2:4] / F(2)
    boxes = np.concatenate([y[..., 0:2] - half, y[..., 0:2] + half], axis=-1)
    m = conf != 0
    cls = np.argmax(y[..., 5:][m], axis=-1).astype(F)[:, None] if m.any() else np.zeros((0, 1), F)
    return np.concatenate([boxes[m], cls], axis=-1)


# --------------------------------------------------------------------------------------------
def get_loss_grad(y_true, y_pred, image_wh, anchors_wh, iou_thresh=0.5, iou_type="iou"):
    """d GetLoss / d y_pred — what tf.GradientTape derives from utils/tf_yolo_utils.py:6-127
    (BCE-with-logits: sigmoid(x) - z; no gradient through the targets or the boolean ignore mask).
    Returns (loss, [grad_l shaped like y_true[l]]).  Checked against fp64 central differences in
    tests/test_oracle_pins.py::test_loss_gradient_matches_finite_differences."""
    image_wh_f = np.asarray(image_wh, F)
    anchors_wh_f = np.asarray(anchors_wh, F)
    loss, _, ign = get_loss(y_true, y_pred, image_wh, anchors_wh, iou_thresh, iou_type, return_ignore=True)
    bsz = np.asarray(y_true[0]).shape[0]
    bf = F(bsz)
    grads = []
    base = 0
    for l in range(3):
        yt = np.asarray(y_true[l], F)
        yp = np.asarray(y_pred[l], F).reshape(yt.shape)
        h, w = yt.shape[1], yt.shape[2]
        n_l = h * w * yt.shape[3]
        ig = ign[:, base:base + n_l].reshape(yt.shape[:4] + (1,)).astype(F)
        base += n_l
        grid = grid_meshgrid(h, w)
        grid_wh = np.array([w, h], dtype=F)
        obj = yt[..., 4:5]
        raw_xy = obj * (yt[..., 0:2] * grid_wh - grid)
        with np.errstate(all="ignore"):
            raw_wh = dm.log((yt[..., 2:4] * image_wh_f[::-1] + F(1e-8)) / anchors_wh_f[l])
        raw_wh = np.where(obj.astype(bool), raw_wh, F(0.0))
        scale = F(2) - yt[..., 2:3] * yt[..., 3:4]
        sig = dm.sigmoid(yp)
        g = np.zeros_like(yp)
        g[..., 0:2] = obj * scale * (sig[..., 0:2] - raw_xy) / bf
        g[..., 2:4] = obj * scale * (yp[..., 2:4] - raw_wh) / bf
        g[..., 4:5] = (sig[..., 4:5] - obj) * (obj + (F(1) - obj) * ig) / bf
        g[..., 5:] = obj * (sig[..., 5:] - yt[..., 5:]) / bf
        grads.append(g)
    return loss, grads


def loss_fp64_fixed_ignore(y_true, y_pred, image_wh, anchors_wh, ignore):
    """The same loss in float64 with the ignore mask held fixed — the differentiable function whose finite
    differences pin get_loss_grad."""
    image_wh_d = np.asarray(image_wh, np.float64)
    total = 0.0
    bsz = np.asarray(y_true[0]).shape[0]
    base = 0
    for l in range(3):
        yt = np.asarray(y_true[l], np.float64)
        yp = np.asarray(y_pred[l], np.float64).reshape(yt.shape)
        h, w = yt.shape[1], yt.shape[2]
        n_l = h * w * yt.shape[3]
        ig = ignore[:, base:base + n_l].reshape(yt.shape[:4] + (1,)).astype(np.float64)
        base += n_l
        grid = grid_meshgrid(h, w).astype(np.float64)
        obj = yt[..., 4:5]
        raw_xy = obj * (yt[..., 0:2] * np.array([w, h], np.float64) - grid)
        with np.errstate(all="ignore"):
            raw_wh = np.log((yt[..., 2:4] * image_wh_d[::-1] + 1e-8) / np.asarray(anchors_wh, np.float64)[l])
        raw_wh = np.where(obj != 0, raw_wh, 0.0)
        scale = 2.0 - yt[..., 2:3] * yt[..., 3:4]
        bce = lambda z, x: np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))
        xy = obj * scale * bce(raw_xy, yp[..., 0:2])
        wh = obj * scale * 0.5 * (raw_wh - yp[..., 2:4]) ** 2
        ob = obj * bce(obj, yp[..., 4:5]) + (1 - obj) * bce(obj, yp[..., 4:5]) * ig
        cl = obj * bce(yt[..., 5:], yp[..., 5:])
        total += (xy.sum() + wh.sum() + ob.sum() + cl.sum()) / bsz
    return total
