"""Golden vectors for the serving path's image pre-processing, produced by the REFERENCE's own code.

utils/image_helper.py needs only cv2 / numpy / PIL, all present in the build container, so this script imports it
straight from /root/reference (unmodified), runs ImageHelper.opencvProportionalResize on seeded synthetic images and
repeats the colour swap / float conversion of views/object_detection.py:56-60.  Small outputs are stored whole, large
ones as sha256 of their bytes (a 416x416 noise image does not compress).  Re-run: python tests/golden/make_golden_letterbox.py
"""
import hashlib
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from letterbox_inputs import CASES, make_image  # noqa: E402

REF = "/root/reference/AIServer/ai_api/ai_models/utils/image_helper.py"
OUT = os.path.join(HERE, "ref_letterbox.npz")


def main():
    import cv2
    spec = importlib.util.spec_from_file_location("ref_image_helper", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    d = {"cv2_version": np.array(cv2.__version__)}
    for i, (name, h, w, size, kind) in enumerate(CASES):
        img_old = make_image(i, h, w, kind)
        bg = (0, 0, 0) if i % 2 == 0 else (128, 128, 128)
        img, _, padding = ref.opencvProportionalResize(img_old, np.int32(size), bg_color=bg)
        predict_img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)          # views/object_detection.py:56
        predict_img = predict_img.astype(np.float32)
        predict_img = predict_img / 255
        predict_img = np.expand_dims(predict_img, 0)
        d[name + "/padding"] = np.int32(padding)
        d[name + "/bg"] = np.int32(bg)
        d[name + "/sha_u8"] = np.array(hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest())
        d[name + "/sha_f32"] = np.array(hashlib.sha256(np.ascontiguousarray(predict_img).tobytes()).hexdigest())
        if img.size <= 130 * 130 * 3:
            d[name + "/u8"] = img
    pts = [[10.0, 20.0], [300.5, 77.25]]
    _, rp, _ = ref.opencvProportionalResize(make_image(0, 480, 640, "noise"), np.int32((96, 96)), points=pts, bg_color=(0, 0, 0))
    d["points_in"], d["points_out"] = np.float64(pts), rp
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, len(CASES), "cases with cv2", cv2.__version__, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
