"""B200-native detection post-network hot path (YOLOv3/v4 + EfficientDet head/loss/eval utilities).

Host-side mirror of the reference's Python call sites (same names, argument order, defaults, layouts) over
the C-ABI library ``libb200det.so`` (hand-written sm_100a CUDA).  Layout mirrors the reference:

    ai_models/utils/tf_iou_utils.py      GetIOU, GetIOUNMS, GetIOUNMSByClasses
    ai_models/utils/tf_yolo_utils.py     GetLoss, GetBoxes, GetNMSBoxes
    ai_models/losses/                    Yolov4Loss, FocalLoss, BoxLoss, ClassFocalLoss
    ai_models/datasets/coco_dataset.py   GetTargets
    ai_models/efficientnet/utils/        Anchors, get_iou, get_nms, get_feat_sizes
    ai_models/efficientnet/efficientdet_net_train.py   get_loss (the _get_loss aggregation)

There is no CPU fallback: importing is cheap, but every compute call needs the built library and a CUDA
device and raises RuntimeError otherwise.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
