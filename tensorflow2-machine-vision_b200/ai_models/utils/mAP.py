"""Drop-in for the reference's utils/mAP.py (Get_mAP_one :114-125), the evaluation step right after NMS in
test_step (yolo_v4/model.py:377, efficientnet/efficientdet_net_train.py:166)."""
import numpy as np
import torch

from ... import _lib, _tensors as T


def Get_mAP_batch(groud_truth, gt_offsets, prediction, pred_offsets, class_num, thresh=0.5):
  '''Per-image mAP for a batch: groud_truth [total,5] (x1,y1,x2,y2,class), prediction [total,6]
  (x1,y1,x2,y2,class,score), offsets [B+1].  Returns a float64 tensor [B].'''
  lib = _lib.load()
  gt = T.to_cuda(groud_truth).reshape(-1, 5)
  pr = T.to_cuda(prediction).reshape(-1, 6)
  go = np.asarray(gt_offsets.cpu() if isinstance(gt_offsets, torch.Tensor) else gt_offsets, dtype=np.int32).reshape(-1)
  po = np.asarray(pred_offsets.cpu() if isinstance(pred_offsets, torch.Tensor) else pred_offsets, dtype=np.int32).reshape(-1)
  B = go.size - 1
  out = torch.zeros((B,), dtype=torch.float64, device=gt.device)
  max_g = int(np.max(np.diff(go))) if B else 0
  max_p = int(np.max(np.diff(po))) if B else 0
  dgo, dpo = T.to_cuda(go, torch.int32), T.to_cuda(po, torch.int32)
  _lib.check(lib.b200_map_per_image(T.ptr(gt), T.ptr(dgo), T.ptr(pr), T.ptr(dpo), B, max_g, max_p, int(class_num),
                                    float(thresh), T.ptr(out), T.stream_ptr()), 'Get_mAP_one')
  return out


def Get_mAP_one(groud_truth, prediction, class_num, thresh=0.5):
  '''mAP of one image (float64 scalar tensor): groud_truth (n,5), prediction (m,6).'''
  n = T.to_cuda(groud_truth).reshape(-1, 5).shape[0]
  m = T.to_cuda(prediction).reshape(-1, 6).shape[0]
  return Get_mAP_batch(groud_truth, [0, n], prediction, [0, m], class_num, thresh)[0]
