"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): NumPy restatement of the serving view's image pre-processing —
ImageHelper.opencvProportionalResize (utils/image_helper.py:293-325, BORDER_CONSTANT) followed by the colour swap and
float conversion of predict (views/object_detection.py:50-62).

The resize itself lives in a third-party dependency of the reference, OpenCV (cv2.resize, INTER_AREA; this image holds
opencv-python 4.13.0).  Its published algorithm for 8-bit images shrinking in both directions is restated here:
  * both scale factors integral  -> block sums in int32, `saturate_cast<uchar>(sum * (1.f / area))` (round half to even);
    the 2x2 case has its own kernel, `(a + b + c + d + 2) >> 2`;
  * otherwise                    -> two fp32 passes over "decimation" taps whose weights come from fp64 interval
    arithmetic (`computeResizeAreaTab`): partial tap, full taps of weight 1/cell, partial tap; products and sums in fp32,
    taps in ascending order, one rounding to uchar at the end.
Pinned: tests/test_letterbox_oracle.py compares it bit for bit with cv2 itself (same image, here and on the GPU box) and
with fixtures produced by the reference's own opencvProportionalResize (tests/golden/make_golden_letterbox.py).
Enlarging in either direction (an input smaller than the network size) is not an area resize in OpenCV: INTER_AREA
falls back to its 8-bit bilinear with "area mode" coefficients (11-bit fixed point, two truncating shifts in the vertical
pass); that path is restated too (linear_coeffs / _resize_bilinear_area_mode).
"""
import math

import numpy as np

F = np.float32
DBL_EPSILON = 2.220446049250313e-16


class UnsupportedResize(ValueError):
    pass


def proportional_size(width, height, new_width, new_height):
    """image_helper.py:297-303 (Python float arithmetic = C double)."""
    if width / height > new_width / new_height:
        rw = new_width
        rh = int((height / width) * rw)
    else:
        rh = new_height
        rw = int((width / height) * rh)
    return rw, rh


def area_taps(ssize, dsize, scale):
    """computeResizeAreaTab: per destination index the (source index, fp32 weight) taps in the order OpenCV applies them."""
    taps = []
    for d in range(dsize):
        fsx1 = d * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        t = []
        if sx1 - fsx1 > 1e-3:
            t.append((sx1 - 1, F((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            t.append((sx, F(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            t.append((sx2, F(min(min(fsx2 - sx2, 1.0), cell) / cell)))
        taps.append(t)
    return taps


def _tap_arrays(taps):
    n = max(len(t) for t in taps)
    si = np.zeros((len(taps), n), np.int64)
    al = np.zeros((len(taps), n), F)
    ok = np.zeros((len(taps), n), bool)
    for d, t in enumerate(taps):
        for k, (s, a) in enumerate(t):
            si[d, k], al[d, k], ok[d, k] = s, a, True
    return si, al, ok


def resize_area(img, dsize):
    """cv2.resize(img, dsize=(w, h), interpolation=cv2.INTER_AREA) for uint8 HxWxC."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    sh, sw = img.shape[:2]
    dw, dh = int(dsize[0]), int(dsize[1])
    if dw < 1 or dh < 1:
        raise UnsupportedResize("empty destination")
    scale_x, scale_y = 1.0 / (dw / sw), 1.0 / (dh / sh)
    if scale_x < 1.0 or scale_y < 1.0:
        return _resize_bilinear_area_mode(img, dw, dh)
    isx, isy = int(round(scale_x)), int(round(scale_y))  # saturate_cast<int>(double) rounds half to even
    if abs(scale_x - isx) < DBL_EPSILON and abs(scale_y - isy) < DBL_EPSILON:
        blk = img[:dh * isy, :dw * isx].reshape(dh, isy, dw, isx, -1).astype(np.int32).sum(axis=(1, 3))
        if isx == 2 and isy == 2 and img.shape[2] in (1, 3, 4):
            return ((blk + 2) >> 2).astype(np.uint8)
        v = blk.astype(F) * F(F(1.0) / F(isx * isy))
        return np.clip(np.rint(v), 0, 255).astype(np.uint8)
    xs, xa, xo = _tap_arrays(area_taps(sw, dw, scale_x))
    ys, ya, yo = _tap_arrays(area_taps(sh, dh, scale_y))
    src = img.astype(F)
    buf = np.zeros((sh, dw, img.shape[2]), F)
    for k in range(xs.shape[1]):  # horizontal pass of every source row: buf += S * alpha, taps in order
        term = (src[:, xs[:, k], :] * xa[None, :, k, None]).astype(F)
        buf = np.where(xo[None, :, k, None], (buf + term).astype(F), buf)
    acc = np.zeros((dh, dw, img.shape[2]), F)
    for k in range(ys.shape[1]):  # vertical pass: sum += beta * buf
        term = (ya[:, k, None, None] * buf[ys[:, k]]).astype(F)
        acc = np.where(yo[:, k, None, None], (acc + term).astype(F), acc)
    return np.clip(np.rint(acc), 0, 255).astype(np.uint8)


def linear_coeffs(ssize, dsize):
    """cv::resize, INTER_AREA outside its true-area domain ("area_mode" of the bilinear branch): per destination index the
    left source index and the two 11-bit fixed-point weights."""
    inv = dsize / ssize
    scale = 1.0 / inv
    idx = np.zeros(dsize, np.int64)
    w = np.zeros((dsize, 2), np.int64)
    for d in range(dsize):
        s = math.floor(d * scale)
        f = F((d + 1) - (s + 1) * inv)
        f = F(0) if f <= 0 else F(f - F(math.floor(f)))
        if s < 0:
            f, s = F(0), 0
        if s >= ssize - 1:
            f, s = F(0), ssize - 1
        idx[d] = s
        w[d, 0] = int(np.clip(np.rint(F(F(1) - f) * F(2048)), -32768, 32767))  # saturate_cast<short>(cbuf * 2048)
        w[d, 1] = int(np.clip(np.rint(f * F(2048)), -32768, 32767))
    return idx, w


def _resize_bilinear_area_mode(img, dw, dh):
    """The 8-bit bilinear of OpenCV (HResizeLinear in int32, VResizeLinear<uchar>: two truncating shifts)."""
    sh, sw = img.shape[:2]
    xi, xw = linear_coeffs(sw, dw)
    yi, yw = linear_coeffs(sh, dh)
    src = img.astype(np.int64)
    x1 = np.minimum(xi + 1, sw - 1)
    hor = src[:, xi, :] * xw[None, :, 0, None] + src[:, x1, :] * xw[None, :, 1, None]   # [sh, dw, C]
    r0 = np.clip(yi, 0, sh - 1)
    r1 = np.clip(yi + 1, 0, sh - 1)
    b0, b1 = yw[:, 0, None, None], yw[:, 1, None, None]
    v = (((b0 * (hor[r0] >> 4)) >> 16) + ((b1 * (hor[r1] >> 4)) >> 16) + 2) >> 2
    return (v & 0xFF).astype(np.uint8)


def proportional_resize(img, size, bg_color=(128, 128, 128)):
    """opencvProportionalResize(img, size, bg_color=..., bg_mode=cv2.BORDER_CONSTANT) -> (image, padding)."""
    height, width = img.shape[:2]
    new_width, new_height = int(size[0]), int(size[1])
    rw, rh = proportional_size(width, height, new_width, new_height)
    small = resize_area(img, (rw, rh))
    top = (new_height - rh) // 2
    bottom = new_height - rh - top
    left = (new_width - rw) // 2
    right = new_width - rw - left
    out = np.empty((new_height, new_width, img.shape[2]), np.uint8)
    out[...] = np.rint(np.asarray(bg_color[:img.shape[2]], np.float64)).clip(0, 255).astype(np.uint8)  # Scalar -> saturate_cast
    out[top:top + rh, left:left + rw] = small
    return out, (top, bottom, left, right)


def predict_preprocess(img_bgr, image_size=(416, 416)):
    """views/object_detection.py:50-62: letterbox on black, BGR->RGB, float32 / 255, batch axis."""
    img, padding = proportional_resize(img_bgr, image_size, bg_color=(0, 0, 0))
    rgb = img[..., ::-1].astype(F)
    return (rgb / 255)[None], padding
