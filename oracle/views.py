"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): NumPy restatement of the box post-processing in the reference's
serving view, views/object_detection.py:70-85.  Pinned against the reference's own lines: tests/golden/make_golden_views.py executes lines 71-85 of the
view verbatim on seeded inputs and tests/test_views_oracle.py holds this file to them (the image half of the view is
oracle/letterbox.py, pinned against OpenCV); NumPy 1.x casting of the reference's era is emulated explicitly (float32 array (op) integer scalar ->
float32), because NumPy >= 2 would promote `float32 * np.int32` to float64."""
import numpy as np

F = np.float32


def restore_predictions(y_boxes, y_classes_id, y_scores, y_classes, y_confidence, image_size, padding, image_size_old):
    y_boxes = np.array(y_boxes, dtype=F).reshape(-1, 4)
    iw, ih = F(image_size[0]), F(image_size[1])
    top, bottom, left, right = [F(v) for v in padding]
    ow, oh = F(image_size_old[0]), F(image_size_old[1])
    # obj:70-71  (y*size - pad) / (size - pad - pad) * size_old, every step rounded to float32
    y_boxes[:, [0, 2]] = ((y_boxes[:, [0, 2]] * iw - left).astype(F) / F(iw - left - right)).astype(F) * ow
    y_boxes[:, [1, 3]] = ((y_boxes[:, [1, 3]] * ih - top).astype(F) / F(ih - top - bottom)).astype(F) * oh
    # obj:73-76 clip
    y_boxes[:, 0][y_boxes[:, 0] < 0] = 0
    y_boxes[:, 1][y_boxes[:, 1] < 0] = 0
    y_boxes[:, 2][y_boxes[:, 2] > ow] = ow
    y_boxes[:, 3][y_boxes[:, 3] > oh] = oh
    # obj:78-84 drop small boxes, int cast
    m = np.logical_and(y_boxes[:, 2] - y_boxes[:, 0] > 2, y_boxes[:, 3] - y_boxes[:, 1] > 2)
    f = lambda a: None if a is None else np.asarray(a)[m]
    return y_boxes[m].astype(np.int32), f(y_classes_id), f(y_scores), f(y_classes), f(y_confidence)
