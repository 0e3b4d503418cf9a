#!/usr/bin/env python
"""How likely is it that TensorFlow's own kernel numerics (Eigen's vectorised exp / logistic / atan / log, which cannot
be run in this image) would flip a DISCRETE result of the hot path — a kept NMS index, an ignore-mask bit, a target
cell — relative to the oracle (detmath) and the NumPy stand-in (libm)?  "Bit-exact versus the reference TF2
implementation" cannot be proven here; this script measures the flip rate instead, three ways:

  A. second back end: the reference's OWN source files run under the TensorFlow stand-in twice — transcendentals from
     NumPy/libm and from PyTorch's CPU kernels (SLEEF; an independent implementation) — and every array of the two runs
     is compared: integer / boolean arrays entry by entry, float arrays in ulps.  (The third implementation, detmath, is
     compared with the libm run by tests/test_reference_emulated.py: discrete outputs identical.)
  B. perturbation: BASELINE config 1 (YOLOv3 416, ~5.3 k candidates, cap 500): every candidate score and/or box
     coordinate is moved by a random -k..+k ulp (k = 1, 2, 4), the NMS is re-run, and the kept index lists are compared.
  C. exposure: how many (record) decisions of the ignore mask sit within a few ulp of the threshold on a config-2
     image, i.e. how many bits COULD flip under a 1-2 ulp change of the metric.

  python tests/golden/tf_numerics_risk.py            (needs /root/reference for part A; writes profiles/r02_tf_numerics_risk.json)
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
F = np.float32


def ulp_diff(a, b):
    a, b = np.asarray(a, F).reshape(-1), np.asarray(b, F).reshape(-1)
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7fffffff), ia)
    ib = np.where(ib < 0, -(ib & 0x7fffffff), ib)
    d = np.abs(ia - ib)
    both_nan = np.isnan(a) & np.isnan(b)
    return np.where(both_nan, 0, d)


def part_a():
    if not os.path.isdir("/root/reference"):
        return {"skipped": "/root/reference not mounted"}
    runs = {}
    with tempfile.TemporaryDirectory() as tmp:
        for backend in ("libm", "torch"):
            out = os.path.join(tmp, backend + ".npz")
            env = dict(os.environ, FAKE_TF_MATH=backend, B200_EMULATED_OUT=out, B200_FORCE_TF_STAND_IN="1")
            r = subprocess.run([sys.executable, os.path.join(HERE, "make_golden_emulated.py")], env=env, capture_output=True, text=True)
            if r.returncode != 0:
                return {"error": r.stderr[-2000:]}
            runs[backend] = dict(np.load(out))
    a, b = runs["libm"], runs["torch"]
    discrete, floats = {}, {}
    for k in sorted(a):
        x, y = a[k], b[k]
        if x.shape != y.shape:
            discrete[k] = {"shape_differs": [list(x.shape), list(y.shape)]}
            continue
        if x.dtype.kind in "iub":
            n = int(np.count_nonzero(x != y))
            discrete[k] = {"entries": int(x.size), "differ": n}
        else:
            d = ulp_diff(x, y)
            floats[k] = {"entries": int(x.size), "differ": int(np.count_nonzero(d)), "max_ulp": int(d.max()) if d.size else 0}
            if k.startswith("gt_") and "_t" in k:   # dense targets: which cells are set is the discrete part
                discrete[k + " (nonzero pattern)"] = {"entries": int(x.size), "differ": int(np.count_nonzero((x != 0) != (y != 0)))}
    # inputs are identical by construction; report only arrays the reference computed
    inputs = [k for k in floats if floats[k]["differ"] == 0]
    return {"backends": ["NumPy/libm", "PyTorch CPU (SLEEF)"],
            "discrete_arrays": len(discrete), "discrete_arrays_with_differences": sorted(k for k, v in discrete.items() if v.get("differ") or v.get("shape_differs")),
            "discrete": discrete,
            "float_arrays_identical": len(inputs),
            "float_arrays_differing": {k: v for k, v in floats.items() if v["differ"]}}


def part_b(trials=12):
    from oracle import yolo as oy
    import emulated_inputs as ei
    heads = ei.yolo_416_heads()
    r = oy.get_nms_boxes_ex(heads[0], heads[1], heads[2], ei.COCO_ANCHORS, (416, 416), 80, 0.5, 0.3, 0.5, "iou")
    cb, cs, cid, base = r["cand_boxes"], r["cand_scores"], r["cand_classes_id"], r["selected"]
    srt = np.sort(cs)
    gaps = ulp_diff(srt[1:], srt[:-1])
    out = {"candidates": int(cs.size), "kept": int(base.size),
           "adjacent_sorted_scores_within_ulp": {str(k): int(np.count_nonzero(gaps <= k)) for k in (0, 1, 2, 4)}}
    rng = np.random.default_rng(20261018 + 99)

    def jitter(x, k):
        i = x.view(np.int32).astype(np.int64) + rng.integers(-k, k + 1, x.shape)
        return i.astype(np.int32).view(F)
    for what in ("scores", "boxes", "scores+boxes"):
        for k in (1, 2, 4):
            changed_runs, changed_slots = 0, 0
            for _ in range(trials):
                s2 = jitter(cs, k) if "scores" in what else cs
                b2 = jitter(cb, k) if "boxes" in what else cb
                sel = oy.get_iou_nms_by_classes(b2, s2, cid, 500, 0.5, "iou")
                same = sel.shape == base.shape and np.array_equal(sel, base)
                changed_runs += 0 if same else 1
                n = min(sel.size, base.size)
                changed_slots += int(np.count_nonzero(sel[:n] != base[:n])) + abs(sel.size - base.size)
            out["%s +-%d ulp" % (what, k)] = {"trials": trials, "runs_with_any_change": changed_runs,
                                              "mean_changed_slots_of_500": changed_slots / trials}
    return out


def part_c():
    from oracle import detmath as dm
    from oracle import yolo as oy
    from tfmv_b200 import synth
    rng = np.random.default_rng(20261018 + 2)
    image, batch = 608, 4
    anc = synth.yolo_anchors().astype(F)
    heads = synth.yolo_heads(rng, batch, image)
    boxes, classes, off = synth.gt_batch(rng, batch, (image, image), max_boxes=100)
    thr = F(0.5)
    near = {1: 0, 2: 0, 4: 0, 16: 0}
    total = 0
    for b in range(batch):
        yt = oy.get_targets(boxes[off[b]:off[b + 1]], classes[off[b]:off[b + 1]], anc, (image, image), 80)
        for l in range(3):
            t = yt[l]
            h, w = t.shape[0], t.shape[1]
            obj = t[..., 4] != 0
            if not obj.any():
                total += t.shape[0] * t.shape[1] * t.shape[2]
                continue
            grid = oy.grid_meshgrid(h, w)
            gwh = np.array([w, h], F)
            yp = heads[l][b].reshape(t.shape)
            pxy = (dm.sigmoid(yp[..., 0:2]) + grid[0]) / gwh
            with np.errstate(all="ignore"):
                pwh = dm.exp(yp[..., 2:4]) * anc[l] / np.array([image, image], F)
            pb = np.concatenate([pxy - pwh / F(2), pxy + pwh / F(2)], -1)
            tb = np.concatenate([t[..., 0:2] - t[..., 2:4] / F(2), t[..., 0:2] + t[..., 2:4] / F(2)], -1)[obj]
            with np.errstate(all="ignore"):
                m = oy.get_iou(pb[..., None, :], tb[None, ...], "ciou")
            best = np.max(m, -1).reshape(-1)
            d = ulp_diff(best, np.full_like(best, thr))
            total += best.size
            for k in near:
                near[k] += int(np.count_nonzero(d <= k))
    return {"records": total, "best_metric_within_ulp_of_threshold": {str(k): v for k, v in near.items()},
            "note": "an ignore bit can flip under a k-ulp change of the metric only for these records"}


def main():
    rep = {"A_second_backend": part_a(), "B_perturbation_config1_nms": part_b(), "C_ignore_mask_exposure_config2": part_c()}
    path = os.path.join(ROOT, "profiles", "r02_tf_numerics_risk.json")
    json.dump(rep, open(path, "w"), indent=1, sort_keys=True)
    a = rep["A_second_backend"]
    print("A:", a.get("discrete_arrays"), "discrete arrays;", "with differences:", a.get("discrete_arrays_with_differences"))
    print("   float arrays differing:", {k: v["max_ulp"] for k, v in (a.get("float_arrays_differing") or {}).items()})
    print("B:", json.dumps(rep["B_perturbation_config1_nms"]))
    print("C:", json.dumps(rep["C_ignore_mask_exposure_config2"]))


if __name__ == "__main__":
    main()
