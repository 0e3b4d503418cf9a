"""Drop-in for the target-assignment part of the reference's datasets/coco_dataset.py
(DataGenerator.GetTargets :185-285).  Image loading / augmentation (:82-183, :287-331) is out of scope."""
import ctypes

import numpy as np
import torch

from ... import _lib, _tensors as T


class DataGenerator():
  def __init__(self, classes_num, anchors, image_wh=(416, 416), **unused):
    '''
    Args:
      anchors: (layers_num, anchors_num, 2) pixels, layer 0 = coarsest head (LoadAnchors order)
      image_wh: (w, h)
    '''
    self.classes_num = int(classes_num)
    self.anchors_wh = np.asarray(anchors)
    self.image_wh = image_wh
    self.layers_hw = [[self.image_wh[1] // i, self.image_wh[0] // i] for i in [32, 16, 8]]

  def GetTargetsBatch(self, classes, boxes, offsets, out=None):
    '''Batched GetTargets: classes [total] int, boxes [total,4] pixel corners, offsets [B+1].
    Returns (target1, target2, target3), each (B, H, W, anchors_num, 5+classes_num).'''
    lib = _lib.load()
    boxes = T.to_cuda(boxes).reshape(-1, 4)
    classes = T.to_cuda(classes, torch.int32).reshape(-1)
    offsets = T.to_cuda(offsets, torch.int32).reshape(-1)
    B = offsets.numel() - 1
    A = self.anchors_wh.shape[1]
    RF = 5 + self.classes_num
    dev = boxes.device
    if out is None:
      out = tuple(torch.empty((B, hw[0], hw[1], A, RF), dtype=torch.float32, device=dev) for hw in self.layers_hw)
    anc = T.host_floats(self.anchors_wh)
    img = T.host_floats(self.image_wh, 2)
    hw = (ctypes.c_int32 * 6)(*[d for l in self.layers_hw for d in l])
    tp = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in out])
    _lib.check(lib.b200_yolo_assign_targets(T.ptr(boxes), T.ptr(classes), T.ptr(offsets), B, boxes.shape[0],
                                            anc.ctypes.data_as(ctypes.c_void_p), A,
                                            img.ctypes.data_as(ctypes.c_void_p), self.classes_num, hw, tp, 1,
                                            T.stream_ptr()), 'GetTargets')
    return out

  def GetTargets(self, img, classes, boxes):
    '''
    One image: boxes (n,4) pixel corners x1,y1,x2,y2, classes (n,) int.
    Returns img, (target1, target2, target3) with targets (H, W, anchors_num, 5+classes_num).
    '''
    n = int(np.prod(boxes.shape[:-1])) if hasattr(boxes, 'shape') else len(boxes)
    t = self.GetTargetsBatch(classes, boxes, np.array([0, n], dtype=np.int32))
    return img, tuple(x[0] for x in t)
