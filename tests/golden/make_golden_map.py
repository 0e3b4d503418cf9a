"""Golden vectors for the mAP row, produced by the REFERENCE's own code.

utils/mAP.py is pure NumPy, so it can run in the build container: this script imports it straight from
/root/reference (restoring the alias np.float = float that NumPy >= 1.24 removed and the reference still uses,
mAP.py:19,25) and freezes Get_mAP_one outputs for the literal example of mAP.py:130-142 and for seeded random cases.
Run from the repo root:  python tests/golden/make_golden_map.py     (needs /root/reference; the .npz is committed)
"""
import importlib.util
import os

import numpy as np

REF = "/root/reference/AIServer/ai_api/ai_models/utils/mAP.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "map_ref.npz")


def load_reference():
    if not hasattr(np, "float"):
        np.float = float  # removed alias the reference still uses
    spec = importlib.util.spec_from_file_location("ref_mAP", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def random_case(rng, n_gt, n_pred, classes, jitter):
    c = rng.random((n_gt, 2)) * 0.8 + 0.1
    wh = np.exp(rng.uniform(np.log(0.05), np.log(0.4), (n_gt, 2)))
    gt = np.concatenate([c - wh / 2, c + wh / 2, rng.integers(0, classes, (n_gt, 1)).astype(np.float64)], -1).astype(np.float32)
    pred = []
    for k in range(n_pred):
        if n_gt and rng.random() < 0.7:
            g = gt[rng.integers(0, n_gt)]
            b = g[:4] + rng.normal(0, jitter, 4)
            cls = g[4] if rng.random() < 0.8 else rng.integers(0, classes)
        else:
            cc = rng.random(2)
            ww = rng.uniform(0.05, 0.3, 2)
            b = np.concatenate([cc - ww / 2, cc + ww / 2])
            cls = rng.integers(0, classes)
        pred.append(np.concatenate([b, [cls, np.round(rng.random(), 3)]]))
    pred = np.asarray(pred, dtype=np.float32).reshape(-1, 6)
    return gt, pred


def main():
    ref = load_reference()
    d = {}
    ex = [
        dict(gt=[[1, 1, 2, 2, 1], [1, 1, 2, 2, 2], [1, 1.3, 2.4, 2, 1], [3, 1, 4, 2, 2]],
             pr=[[1.1, 1, 2.1, 2.2, 1, 0.8], [1.2, 1.2, 2.2, 2.2, 2, 0.7], [1.1, 1.3, 2.4, 2.1, 1, 0.6], [1.1, 1.1, 2.1, 2.1, 1, 0.9]]),
        dict(gt=[[1, 1, 2, 2, 1], [1, 1, 2, 2, 2], [1, 1.3, 2.4, 2, 1], [3, 1, 4, 2, 2], [3, 1, 4, 2, 0]],
             pr=[[1.1, 1, 2.1, 2.2, 1, 0.8], [1.2, 1.2, 2.2, 2.2, 2, 0.7], [1.1, 1.3, 2.4, 2.1, 1, 0.7], [1.1, 1.1, 2.1, 2.1, 1, 0.6]]),
    ]
    n = 0
    for e in ex:  # the literal fixture of mAP.py:130-142, image by image (Get_mAP_one is the per-image entry test_step uses)
        gt, pr = np.asarray(e["gt"], np.float32), np.asarray(e["pr"], np.float32)
        d["gt%d" % n], d["pr%d" % n], d["classes%d" % n] = gt, pr, np.int32(3)
        d["map%d" % n] = np.float64(ref.Get_mAP_one(gt, pr, 3, 0.5))
        n += 1
    rng = np.random.default_rng(77)
    for (n_gt, n_pred, classes, jitter) in [(1, 1, 2, 0.01), (5, 9, 3, 0.02), (20, 60, 8, 0.03), (40, 200, 80, 0.02),
                                            (0, 5, 4, 0.02), (6, 0, 4, 0.02), (100, 500, 80, 0.03), (30, 100, 5, 0.08)]:
        gt, pr = random_case(rng, n_gt, n_pred, classes, jitter)
        d["gt%d" % n], d["pr%d" % n], d["classes%d" % n] = gt, pr, np.int32(classes)
        d["map%d" % n] = np.float64(ref.Get_mAP_one(gt, pr, classes, 0.5)) if n_pred else np.float64(0.0)
        n += 1
    d["count"] = np.int32(n)
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, n, "cases; fixture mAP:", float(d["map0"]), float(d["map1"]))


if __name__ == "__main__":
    main()
