"""Drop-in for the target-assignment part of the reference's datasets/coco_dataset.py
(DataGenerator.GetTargets :185-285).  Image loading / augmentation (:82-183, :287-331) is out of scope."""
import ctypes

import numpy as np
import torch

from ... import _lib, _tensors as T


class TargetBuffers():
  '''Dense y_true tensors that are reused step after step.  The first GetTargetsBatch into them zero-fills the
  tensors; later calls zero only the records written by the previous call (b200_yolo_reset_targets), which turns
  the 7.7 MB/image dense re-fill at 608x608 into a few kilobytes.  The tensors handed back are the same objects
  every step and are complete dense targets (bit-identical to a fresh zero-filled assignment); callers must not
  write into them.'''
  def __init__(self):
    self.targets = None       # tuple of 3 dense tensors
    self.prev_boxes = None    # device copy of the previous call's boxes [cap,4]
    self.prev_offsets = None  # device copy of the previous call's offsets [B+1]
    self.prev_total = 0
    self.key = None
    self.needs_fill = True

  def invalidate(self):
    '''Forget the previous ground truth: the next call does a dense zero-fill (use after writing into the tensors).'''
    self.needs_fill = True


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

  def GetTargetsBatch(self, classes, boxes, offsets, out=None, buffers=None):
    '''Batched GetTargets: classes [total] int, boxes [total,4] pixel corners, offsets [B+1].
    Returns (target1, target2, target3), each (B, H, W, anchors_num, 5+classes_num).
    buffers: optional TargetBuffers — persistent output tensors with sparse reset instead of the dense zero-fill.'''
    lib = _lib.load()
    boxes = T.to_cuda(boxes).reshape(-1, 4)
    classes = T.to_cuda(classes, torch.int32).reshape(-1)
    offsets = T.to_cuda(offsets, torch.int32).reshape(-1)
    B = offsets.numel() - 1
    A = self.anchors_wh.shape[1]
    RF = 5 + self.classes_num
    dev = boxes.device
    if buffers is not None:
      return self._targets_into_buffers(lib, buffers, classes, boxes, offsets, B, A, RF)
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

  def _targets_into_buffers(self, lib, buf, classes, boxes, offsets, B, A, RF):
    dev = boxes.device
    anc = T.host_floats(self.anchors_wh)
    img = T.host_floats(self.image_wh, 2)
    hw = (ctypes.c_int32 * 6)(*[d for l in self.layers_hw for d in l])
    total = boxes.shape[0]
    key = (B, A, RF, tuple(map(tuple, self.layers_hw)), str(dev), anc.tobytes(), img.tobytes())
    if buf.key != key:
      buf.needs_fill = True
      buf.targets = tuple(torch.empty((B, l[0], l[1], A, RF), dtype=torch.float32, device=dev) for l in self.layers_hw)
      buf.prev_offsets = torch.zeros(B + 1, dtype=torch.int32, device=dev)
      buf.prev_boxes = torch.zeros((max(total, 1), 4), dtype=torch.float32, device=dev)
      buf.key = key
    fresh = buf.needs_fill
    buf.needs_fill = False
    tp = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in buf.targets])
    if not fresh and buf.prev_total > 0:
      _lib.check(lib.b200_yolo_reset_targets(T.ptr(buf.prev_boxes), T.ptr(buf.prev_offsets), B, buf.prev_total,
                                             anc.ctypes.data_as(ctypes.c_void_p), A, img.ctypes.data_as(ctypes.c_void_p),
                                             self.classes_num, hw, tp, T.stream_ptr()), 'GetTargets')
    _lib.check(lib.b200_yolo_assign_targets(T.ptr(boxes), T.ptr(classes), T.ptr(offsets), B, total,
                                            anc.ctypes.data_as(ctypes.c_void_p), A,
                                            img.ctypes.data_as(ctypes.c_void_p), self.classes_num, hw, tp, 1 if fresh else 0,
                                            T.stream_ptr()), 'GetTargets')
    # remember this call's ground truth (own copies: the caller may reuse its arrays)
    if buf.prev_boxes.shape[0] < total:
      buf.prev_boxes = torch.zeros((total, 4), dtype=torch.float32, device=dev)
    if total > 0:
      buf.prev_boxes[:total].copy_(boxes)
    buf.prev_offsets.copy_(offsets)
    buf.prev_total = total
    return buf.targets

  def GetTargets(self, img, classes, boxes):
    '''
    One image: boxes (n,4) pixel corners x1,y1,x2,y2, classes (n,) int.
    Returns img, (target1, target2, target3) with targets (H, W, anchors_num, 5+classes_num).
    '''
    n = int(np.prod(boxes.shape[:-1])) if hasattr(boxes, 'shape') else len(boxes)
    t = self.GetTargetsBatch(classes, boxes, np.array([0, n], dtype=np.int32))
    return img, tuple(x[0] for x in t)
