"""CPU oracle for the detection hot path — TEST INFRASTRUCTURE ONLY.

A NumPy (fp32, op-by-op, same operation order) restatement of the reference's TensorFlow code for
YOLOv3/v4 head decode, IoU families, target assignment, yolo_loss, NMS, and the EfficientDet
anchor / focal / box-loss utilities.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package; the product
package (``tensorflow2-machine-vision_b200``) never does and fails loudly without its CUDA library.

PARITY STATUS: **parity unpinned** for numeric values.  The reference is pure Python on top of
TensorFlow 2.x (unpinned, not vendored, not installable here: ``import tensorflow`` fails and there
is no network), it has no C/C++ path to compile into ``oracle/_ref`` and its only two unit tests
(yolo_v3/unit_test/grid_test.py, loss_test.py) assert *relations* (grid layout equality; GetLoss-copy
== Yolov4Loss), not values.  What IS pinned (tests/test_oracle_*.py):
  * both relations above, re-checked on this oracle;
  * hand-derived known answers for the literal inputs the reference ships
    (efficientnet/utils/iou.py:104-111, tests/test_anchors.py:10-15);
  * TF op semantics listed in SURVEY.md §8a (argsort ties, argmax first-max, scatter_nd duplicate
    sums, floor-div, divide_no_nan, boolean_mask order, sigmoid_cross_entropy formula).
Transcendentals go through the deterministic fp32 header shared with the kernels
(csrc/detmath.h, <= 3 ulp from correctly rounded; oracle/DETMATH_REPORT.md), so discrete outputs
can be compared bit-for-bit with the GPU.
"""
