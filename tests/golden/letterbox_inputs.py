"""Seeded synthetic images for the letterbox goldens (shared by the generator and the tests)."""
import numpy as np

# (name, source height, source width, target (w, h), kind)
CASES = [
    ("noise_640x480_to_96x96", 480, 640, (96, 96), "noise"),
    ("noise_333x777_to_64x80", 777, 333, (64, 80), "noise"),
    ("noise_exact_3x", 288, 384, (128, 96), "noise"),          # both scales exactly 3: block path
    ("noise_exact_2x", 256, 256, (128, 128), "noise"),         # the 2x2 path
    ("noise_same_size", 96, 128, (128, 96), "noise"),          # scale 1
    ("binary_1001x701_to_128x128", 701, 1001, (128, 128), "binary"),
    ("smooth_1920x1080_to_416x416", 1080, 1920, (416, 416), "smooth"),
    ("noise_1920x1080_to_416x416", 1080, 1920, (416, 416), "noise"),
    ("noise_4032x3024_to_416x416", 3024, 4032, (416, 416), "noise"),
    ("noise_tall_500x2000_to_416x416", 2000, 500, (416, 416), "noise"),
    ("noise_wide_to_608", 900, 1600, (608, 608), "noise"),
    ("noise_one_row_short", 417, 416, (416, 416), "noise"),    # scale_x == 1 after int(), scale_y slightly above 1
    # inputs smaller than the target: INTER_AREA becomes OpenCV's 8-bit bilinear ("area mode" coefficients)
    ("small_100x80_to_128x128", 80, 100, (128, 128), "noise"),
    ("small_binary_37x91_to_128x96", 91, 37, (128, 96), "binary"),
    ("small_320x240_to_416x416", 240, 320, (416, 416), "noise"),
    ("small_mixed_415x10_to_416x416", 10, 415, (416, 416), "noise"),   # grows along x, scale 1 along y
    ("small_1x1_to_64x64", 1, 1, (64, 64), "noise"),
]


def make_image(i, h, w, kind):
    rng = np.random.default_rng(20261018 + i)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == "binary":  # only 0 / 255: sums land on .5 often
        return (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([(x * 255 // max(w - 1, 1)), (y * 255 // max(h - 1, 1)), ((x + y) % 256)], -1).astype(np.uint8)
    img[h // 4:h // 2, w // 3:w // 2] = (200, 30, 90)
    return img
