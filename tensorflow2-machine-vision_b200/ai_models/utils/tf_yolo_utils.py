"""Drop-in for the reference's utils/tf_yolo_utils.py: GetLoss (:6-127), GetBoxes (:129-167), GetNMSBoxes (:169-269).

Same names, positional order, defaults and NHWC layouts; tensors may be torch CUDA tensors, anything exporting
DLPack (tf.experimental.dlpack.to_dlpack on a TF box), or host arrays (copied).  Results are torch CUDA tensors.
"""
import ctypes

import numpy as np
import torch

from ... import _lib, _tensors as T

NMS_MAX_OUTPUT = 500  # hard-coded in the reference, tf_yolo_utils.py:259


def GetBoxes(y, anchors_wh, classes_num):
  '''Head offsets -> boxes in [0,1].

  Args:
    y: (batch, H, W, anchors_num, 5+classes_num)
    anchors_wh: (anchors_num, 2), already divided by image_wh
  Returns:
    boxes (N,4) x1,y1,x2,y2; confidence (N,1); classes (N,classes_num) — rows with x2<=x1 or y2<=y1 removed,
    batch dimension flattened, row-major order (reference :163-166).
  '''
  lib = _lib.load()
  y = T.to_cuda(y)
  B, H, W, A = y.shape[0], y.shape[1], y.shape[2], y.shape[3]
  C = int(classes_num)
  if y.shape[4] != 5 + C:
    raise ValueError('last dimension %d != 5 + classes_num %d' % (y.shape[4], C))
  anc = T.to_cuda(anchors_wh).reshape(-1)
  if anc.numel() != 2 * A:
    raise ValueError('anchors_wh must be (anchors_num, 2)')
  n = B * H * W * A
  boxes = torch.empty((n, 4), dtype=torch.float32, device=y.device)
  conf = torch.empty((n, 1), dtype=torch.float32, device=y.device)
  classes = torch.empty((n, C), dtype=torch.float32, device=y.device)
  valid = torch.empty((n,), dtype=torch.uint8, device=y.device)
  _lib.check(lib.b200_yolo_decode_dense(T.ptr(y), B, H, W, A, C, T.ptr(anc), T.ptr(boxes), T.ptr(conf),
                                        T.ptr(classes), T.ptr(valid), T.stream_ptr()), 'GetBoxes')
  keep = valid.bool()  # order-preserving row selection (device-side plumbing, like tf.boolean_mask)
  return boxes[keep], conf[keep], classes[keep]


def _levels(y1, y2, y3, anchors_wh):
  anc = T.host_floats(anchors_wh)
  if anc.size % 6 != 0:
    raise ValueError('anchors_wh must be (3, anchors_num, 2)')
  A = anc.size // 6
  heads = []
  for y in (y1, y2, y3):
    t = T.to_cuda(y)
    if t.dim() == 5:
      t = t.reshape(t.shape[0], t.shape[1], t.shape[2], -1)
    heads.append(t)
  B = heads[0].shape[0]
  ch = heads[0].shape[3]
  if ch % A != 0:
    raise ValueError('channel count %d is not a multiple of anchors_num %d' % (ch, A))
  for t in heads:
    if t.shape[0] != B or t.shape[3] != ch:
      raise ValueError('the three heads disagree on batch or channel size')
  return heads, anc, A, B, ch // A


def GetNMSBoxesBatch(y1, y2, y3, anchors_wh, image_wh, classes_num,
  confidence_thresh=0.5, scores_thresh=0.3, iou_thresh=0.5, iou_type='iou',
  max_output_size=NMS_MAX_OUTPUT, with_classes=True, with_indices=False):
  '''Batched form of GetNMSBoxes: the reference's B == 1 semantics applied to every image.

  Returns a dict of padded tensors [B, max_output_size, ...] plus `count` [B] (int32):
  boxes, classes_id, scores, classes (optional), confidence, and with_indices -> sel_idx / sel_anchor.
  '''
  assert iou_type in ['iou','diou','ciou']
  lib = _lib.load()
  heads, anc, A, B, RF = _levels(y1, y2, y3, anchors_wh)
  C = int(classes_num)
  if RF != 5 + C:
    raise ValueError('head channels %d != anchors_num*(5+classes_num)' % (RF * A))
  img = T.host_floats(image_wh, 2)
  dev = heads[0].device
  K = int(max_output_size)
  hw = (ctypes.c_int32 * 6)(*[d for t in heads for d in (t.shape[1], t.shape[2])])
  hp = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in heads])
  out = {
    'boxes': torch.empty((B, K, 4), dtype=torch.float32, device=dev),
    'classes_id': torch.empty((B, K), dtype=torch.int32, device=dev),
    'scores': torch.empty((B, K), dtype=torch.float32, device=dev),
    'confidence': torch.empty((B, K, 1), dtype=torch.float32, device=dev),
    'count': torch.zeros((B,), dtype=torch.int32, device=dev),
  }
  if with_classes:
    out['classes'] = torch.empty((B, K, C), dtype=torch.float32, device=dev)
  if with_indices:
    out['sel_idx'] = torch.empty((B, K), dtype=torch.int32, device=dev)
    out['sel_anchor'] = torch.empty((B, K), dtype=torch.int32, device=dev)
  ws_bytes = lib.b200_yolo_decode_nms_workspace_bytes(hw, B, A, K)
  ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
  _lib.check(lib.b200_yolo_decode_nms(
    hp, hw, B, A, C, anc.ctypes.data_as(ctypes.c_void_p), img.ctypes.data_as(ctypes.c_void_p),
    float(confidence_thresh), float(scores_thresh), float(iou_thresh), _lib.METRIC_YOLO[iou_type], K,
    T.ptr(out['boxes']), T.ptr(out['classes_id']), T.ptr(out['scores']), T.ptr(out.get('classes')),
    T.ptr(out['confidence']), T.ptr(out.get('sel_idx')), T.ptr(out.get('sel_anchor')), T.ptr(out['count']),
    T.ptr(ws), ws_bytes, T.stream_ptr()), 'GetNMSBoxes')
  return out


def GetNMSBoxes(y1, y2, y3, anchors_wh, image_wh, classes_num,
  confidence_thresh=0.5, scores_thresh=0.3, iou_thresh=0.5, iou_type='iou'):
  '''Decode the three heads, threshold, per-class NMS (cap 500), gather.

  Returns (selected_boxes (K,4), selected_classes_id (K,) int32, selected_scores (K,),
  selected_classes (K,classes_num), selected_confidence (K,1)) exactly as the reference for batch size 1.
  For batch size > 1 (where the reference would run one NMS over the flattened batch) the per-image results
  are concatenated in image order; use GetNMSBoxesBatch to keep them apart.
  '''
  r = GetNMSBoxesBatch(y1, y2, y3, anchors_wh, image_wh, classes_num, confidence_thresh, scores_thresh,
                       iou_thresh, iou_type)
  cnt = r['count'].cpu().tolist()
  pick = lambda t: torch.cat([t[b, :cnt[b]] for b in range(len(cnt))], dim=0) if len(cnt) != 1 else t[0, :cnt[0]]
  return pick(r['boxes']), pick(r['classes_id']), pick(r['scores']), pick(r['classes']), pick(r['confidence'])
