"""CPU oracle for the detection hot path — TEST INFRASTRUCTURE ONLY.

A NumPy (fp32, op-by-op, same operation order) restatement of the reference's TensorFlow code for
YOLOv3/v4 head decode, IoU families, target assignment, yolo_loss, NMS, and the EfficientDet
anchor / focal / box-loss utilities.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package; the product
package (``tensorflow2-machine-vision_b200``) never does and fails loudly without its CUDA library.

PARITY STATUS: **parity unpinned against TensorFlow's kernels; pinned against the reference's own source.**
The reference is pure Python on top of TensorFlow 2.x (unpinned, not vendored, not installable here:
``import tensorflow`` fails and there is no network), it has no C/C++ path to compile into ``oracle/_ref``
and its only two unit tests (yolo_v3/unit_test/grid_test.py, loss_test.py) assert *relations* (grid layout
equality; GetLoss-copy == Yolov4Loss), not values.  What IS pinned:
  * the reference's OWN source files for this path (tf_iou_utils.py, tf_yolo_utils.py GetLoss / GetBoxes /
    GetNMSBoxes, datasets/coco_dataset.py GetTargets, efficientnet/utils/{iou,nms,anchors}.py,
    losses/{focal_loss,box_loss,class_loss,yolo_loss}.py, efficientdet_net_train.py _get_loss, yolo_v4/model.py
    GetGroudTruth), imported UNMODIFIED from /root/reference and executed under a
    NumPy stand-in for the ~60 TensorFlow ops they use (tests/golden/fake_tf, tests/golden/make_golden_emulated.py
    -> tests/golden/ref_emulated.npz): tests/test_reference_emulated.py holds this oracle to those outputs —
    NMS indices, class ids, masks, one-hot rows, anchors and dense targets identical, floating-point results to a few
    ulp (libm vs detmath).  That pins control flow, operation order, broadcasting and index conventions; it cannot
    pin TensorFlow's kernel numerics;
  * both unit-test relations above, re-checked on this oracle (tests/test_oracle_pins.py);
  * hand-derived known answers for the literal inputs the reference ships
    (efficientnet/utils/iou.py:104-111, tests/test_anchors.py:10-15) — which the reference's code reproduces under
    the stand-in as well;
  * utils/mAP.py (pure NumPy) run directly (tests/golden/make_golden_map.py);
  * utils/image_helper.py opencvProportionalResize (needs only OpenCV, present here) run directly
    (tests/golden/make_golden_letterbox.py), and oracle/letterbox.py against cv2.resize itself (PINNED row);
  * TF op semantics listed in SURVEY.md §8a (argsort ties, argmax first-max, scatter_nd duplicate
    sums, floor-div, divide_no_nan, boolean_mask order, sigmoid_cross_entropy formula).
Transcendentals go through the deterministic fp32 header shared with the kernels
(csrc/detmath.h, <= 3 ulp from correctly rounded; oracle/DETMATH_REPORT.md), so discrete outputs
can be compared bit-for-bit with the GPU.
"""
