"""One-shot parity of the CUDA path against the committed golden fixtures (yolo_64.npz, effdet_64.npz), per BASELINE
config, for bench.py's `configs[*].parity` and for tests/test_golden.py.  Uses only the product package and the .npz
files — neither the oracle nor the reference tree is consulted at run time."""
import os

import numpy as np

GOLD = os.path.dirname(os.path.abspath(__file__))
F = np.float32


def _dense(shape, idx, val):
    t = np.zeros(shape, F)
    t[tuple(idx.T)] = val
    return t


def _yolo():
    g = np.load(os.path.join(GOLD, "yolo_64.npz"))
    heads = [g["h0"], g["h1"], g["h2"]]
    batch = heads[0].shape[0]
    y_true = [_dense((batch, gr, gr, 3, 85), g["t%d_idx" % l], g["t%d_val" % l]) for l, gr in enumerate((2, 4, 8))]
    return g, heads, y_true


def _nms_ok(r, g, batch):
    ok = True
    for b in range(batch):
        k = int(r["count"][b])
        ok &= r["sel_idx"][b, :k].cpu().tolist() == g["nms%d_selected" % b].tolist()
        ok &= np.array_equal(r["boxes"][b, :k].cpu().numpy(), g["nms%d_boxes" % b])
        ok &= np.array_equal(r["scores"][b, :k].cpu().numpy(), g["nms%d_scores" % b])
        ok &= r["classes_id"][b, :k].cpu().tolist() == g["nms%d_classes_id" % b].tolist()
    return bool(ok)


def yolo_decode_nms(dev):
    """c1 (and the decode half of c5): GetNMSBoxes — kept indices, boxes, scores, class ids identical to the fixture."""
    import torch
    from tfmv_b200.ai_models.utils.tf_yolo_utils import GetNMSBoxesBatch
    g, heads, _ = _yolo()
    anc, image = g["anchors"], int(g["image"])
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    r = GetNMSBoxesBatch(*[d(h) for h in heads], anc, (image, image), 80, 0.5, 0.3, 0.5, "diou", with_indices=True)
    return {"ok": _nms_ok(r, g, heads[0].shape[0]), "checked": "NMS indices / boxes / scores / class ids bit-exact (yolo_64.npz)"}


def yolo_targets_loss(dev):
    """c2: GetTargets bit-exact, ignore mask bit-exact, loss within 1e-4 relative."""
    import torch
    from tfmv_b200.ai_models.datasets.coco_dataset import DataGenerator
    from tfmv_b200.ai_models.utils.tf_yolo_utils import _loss_call
    g, heads, y_true = _yolo()
    anc, image = g["anchors"], int(g["image"])
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    gen = DataGenerator(80, anc / F(image), (image, image))
    gen.layers_hw = [[2, 2], [4, 4], [8, 8]]
    t = gen.GetTargetsBatch(d(g["gt_classes"]), d(g["gt_boxes"]), d(g["gt_offsets"]))
    ok = all(np.array_equal(t[l].cpu().numpy(), y_true[l]) for l in range(3))
    worst = 0.0
    for it in ("iou", "ciou"):
        ign = torch.zeros((heads[0].shape[0], 84 * 3), dtype=torch.uint8, device=dev)
        loss = _loss_call([d(x) for x in y_true], [d(h) for h in heads], (image, image), anc / F(image), 0.5, it, 0, ignore_out=ign)
        ok &= np.array_equal(np.packbits(ign.cpu().numpy(), axis=1), g["ignore_" + it])
        worst = max(worst, abs(float(loss) - float(g["loss_" + it])) / abs(float(g["loss_" + it])))
    ok &= worst <= 1e-4
    return {"ok": bool(ok), "loss_rel_err": worst, "checked": "targets and ignore mask bit-exact, loss <= 1e-4 relative (yolo_64.npz)"}


def yolo_loss_and_nms(dev):
    """c5: the loss + decode + NMS step on one y_pred (the fused entry point when the package has one)."""
    import torch
    from tfmv_b200.ai_models.utils import tf_yolo_utils as tyu
    g, heads, y_true = _yolo()
    anc, image = g["anchors"], int(g["image"])
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    fused = getattr(tyu, "LossAndNMSBoxesBatch", None)
    yt, hp = [d(x) for x in y_true], [d(h) for h in heads]
    if fused is not None:
        loss, r = fused(yt, hp, (image, image), anc, 80, 0.5, "ciou", 0.5, 0.3, 0.5, "diou", with_indices=True,
                        loss_anchors_wh=anc / F(image))
    else:
        loss = tyu._loss_call(yt, hp, (image, image), anc / F(image), 0.5, "ciou", 0)
        r = tyu.GetNMSBoxesBatch(*hp, anc, (image, image), 80, 0.5, 0.3, 0.5, "diou", with_indices=True)
    err = abs(float(loss) - float(g["loss_ciou"])) / abs(float(g["loss_ciou"]))
    return {"ok": bool(_nms_ok(r, g, heads[0].shape[0]) and err <= 1e-4), "loss_rel_err": err, "fused_entry": fused is not None,
            "checked": "NMS outputs bit-exact and loss <= 1e-4 relative on the same y_pred (yolo_64.npz)"}


def _effdet_anchors():
    from tfmv_b200.ai_models.efficientnet.utils.anchors import Anchors
    return Anchors(3, 5, (64, 64), 3, [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)], 4.0)


def _post_ok(r, g, n):
    ok = True
    for b in range(n):
        k = int(r["count"][b])
        ok &= r["sel_idx"][b, :k].cpu().tolist() == g["post%d_selected" % b].tolist()
        ok &= np.array_equal(r["boxes"][b, :k].cpu().numpy(), g["post%d_boxes" % b])
        ok &= np.array_equal(r["scores"][b, :k].cpu().numpy(), g["post%d_scores" % b])
        ok &= r["classes_id"][b, :k].cpu().tolist() == g["post%d_classes_id" % b].tolist()
    return bool(ok)


def effdet_decode_post(dev):
    """c4: convert_outputs_boxes + convert_outputs_one (fused entry when present)."""
    import torch
    g = np.load(os.path.join(GOLD, "effdet_64.npz"))
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    a = _effdet_anchors()
    L = 3
    pb, pc = [d(g["pb%d" % l]) for l in range(L)], [d(g["pc%d" % l]) for l in range(L)]
    if hasattr(a, "decode_and_postprocess"):
        dec, r = a.decode_and_postprocess(pb, pc, with_indices=True)
    else:
        dec = a.convert_outputs_boxes(pb)
        r = a.convert_outputs_batch(dec, pc, with_indices=True)
    ok = all(np.array_equal(dec[l].cpu().numpy(), g["dec%d" % l]) for l in range(L)) and _post_ok(r, g, 2)
    return {"ok": bool(ok), "fused_entry": hasattr(a, "decode_and_postprocess"),
            "checked": "decoded boxes, NMS indices / boxes / scores / class ids bit-exact (effdet_64.npz)"}


def effdet_eval_step(dev):
    """c3: focal + box loss, decode, post-process (the fused eval step when present)."""
    import torch
    from tfmv_b200.ai_models.efficientnet import efficientdet_net_train as edt
    g = np.load(os.path.join(GOLD, "effdet_64.npz"))
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    a = _effdet_anchors()
    L, C = 3, int(g["C"])
    tb, tc, tm = a.generate_targets_batch(d(g["gt_boxes"]), d(g["gt_classes"]), d(g["gt_offsets"]), C)
    ok = all(np.array_equal(tb[l].cpu().numpy(), g["tb%d" % l]) and np.array_equal(tm[l].cpu().numpy(), g["tm%d" % l]) and
             np.array_equal(tc[l].argmax(-1).cpu().numpy(), g["tcid%d" % l]) for l in range(L))
    pb, pc = [d(g["pb%d" % l]) for l in range(L)], [d(g["pc%d" % l]) for l in range(L)]
    if hasattr(a, "eval_step"):
        loss, dec, r = a.eval_step(tb, tc, tm, pb, pc, with_indices=True)
    else:
        loss = edt.get_loss(tb, tc, tm, pb, pc)
        dec = a.convert_outputs_boxes(pb)
        r = a.convert_outputs_batch(dec, pc, with_indices=True)
    err = abs(float(loss) - float(g["loss"])) / abs(float(g["loss"]))
    ok = ok and all(np.array_equal(dec[l].cpu().numpy(), g["dec%d" % l]) for l in range(L)) and _post_ok(r, g, 2) and err <= 1e-4
    return {"ok": bool(ok), "loss_rel_err": err, "fused_entry": hasattr(a, "eval_step"),
            "checked": "targets, decoded boxes and NMS outputs bit-exact, loss <= 1e-4 relative (effdet_64.npz)"}


def run_all(dev):
    out = {}
    for name, fn in (("c1", yolo_decode_nms), ("c2", yolo_targets_loss), ("c3", effdet_eval_step), ("c4", effdet_decode_post),
                     ("c5", yolo_loss_and_nms)):
        try:
            out[name] = fn(dev)
        except Exception as e:  # noqa: BLE001
            out[name] = {"ok": False, "error": repr(e)}
    return out
