"""Drop-in for the loss aggregation of the reference's efficientnet/efficientdet_net_train.py
(EfficientDetNetTrain._get_loss :41-52).  The L2-regularisation term (:21-28,42) is over model weights and stays
in the TF graph; add it to the value returned here."""
import ctypes

import numpy as np
import torch

from ... import _lib, _tensors as T
from ..losses.focal_loss import _partial_sums


def get_loss(y_true_boxes, y_true_classes, y_true_masks, y_pred_boxes, y_pred_classes, alpha=0.25, gamma=1.5,
             delta=0.1, global_batch_scale=1, group=None, return_parts=False, with_grad=False, exchange=None):
  '''sum over levels of (50 * box_loss + focal_loss) with num_positives = sum(masks) + 1.

  Data parallel: every rank passes its shard; the 2L+1 fp64 partial sums are all-reduced once before the
  normalisation — inside the finalize kernel over NVLink peer mailboxes (exchange=runtime.PeerExchange), by the
  library's ncclAllReduce (runtime.NcclExchange), or by torch.distributed on `group` (exchange=None).  `global_batch_scale` = world size when the per-level element count of the Keras mean must refer
  to the global batch.
  '''
  lib = _lib.load()
  sums, numel = _partial_sums(list(y_true_boxes), list(y_true_classes), list(y_true_masks), list(y_pred_boxes),
                              list(y_pred_classes), alpha, gamma, delta, 0.0)
  import torch.distributed as dist
  peer = exchange is not None and hasattr(exchange, 'mailboxes')
  if exchange is not None:
    # runtime.NcclExchange: ncclAllReduce issued by the library; runtime.PeerExchange: summed inside the finalize kernel
    if not peer:
      exchange.allreduce_(sums)
    global_batch_scale = exchange.world
  elif dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    global_batch_scale = dist.get_world_size(group)
  L = len(numel)
  nm = (ctypes.c_double * L)(*[n * global_batch_scale for n in numel])
  parts = torch.empty((L, 2), dtype=torch.float32, device=sums.device)
  loss = torch.empty((), dtype=torch.float32, device=sums.device)
  npos = torch.empty((), dtype=torch.float32, device=sums.device)
  if peer:
    rank, world, boxes_ = exchange.args()
    _lib.check(lib.b200_focal_box_finalize_dp(L, T.ptr(sums), nm, T.ptr(parts), T.ptr(loss), T.ptr(npos), rank, world, boxes_,
                                              T.stream_ptr()), '_get_loss (data parallel)')
  else:
    _lib.check(lib.b200_focal_box_finalize(L, T.ptr(sums), nm, T.ptr(parts), T.ptr(loss), T.ptr(npos), T.stream_ptr()),
               '_get_loss')
  if with_grad:
    tb = [T.to_cuda(t) for t in y_true_boxes]
    # class targets: one-hot float tensors, or integer class ids (generate_targets_batch(class_index=True)) — the same
    # detection as the forward pass uses
    indexed = any(not torch.is_floating_point(torch.as_tensor(t)) for t in y_true_classes)
    tc = [T.to_cuda(t, torch.int32) if indexed else T.to_cuda(t) for t in y_true_classes]
    pb = [T.to_cuda(t) for t in y_pred_boxes]
    pc = [T.to_cuda(t) for t in y_pred_classes]
    gb = [torch.empty_like(t) for t in pb]
    gc = [torch.empty_like(t) for t in pc]
    C = pc[0].shape[-1]
    for l in range(L):
      if tc[l].numel() * (C if indexed else 1) != pc[l].numel():
        raise ValueError('class target / output shapes differ at level %d' % l)
    anc = (ctypes.c_ulonglong * L)(*[t.numel() // C for t in pc])
    arr = lambda ts: (ctypes.c_void_p * L)(*[t.data_ptr() for t in ts])
    grad = lib.b200_focal_box_grad_indexed if indexed else lib.b200_focal_box_grad
    _lib.check(grad(L, anc, C, arr(tb), arr(tc), arr(pb), arr(pc), float(alpha), float(gamma),
                    float(delta), 0.0, T.ptr(sums), nm, arr(gb), arr(gc), T.stream_ptr()),
               '_get_loss backward')
    return loss, tuple(gb), tuple(gc)
  return (loss, parts, npos) if return_parts else loss


def get_loss_and_grad(y_true_boxes, y_true_classes, y_true_masks, y_pred_boxes, y_pred_classes, alpha=0.25, gamma=1.5,
                      delta=0.1):
  '''_get_loss plus (d loss / d y_pred_boxes[l], d loss / d y_pred_classes[l]) for an upstream gradient of 1.'''
  return get_loss(y_true_boxes, y_true_classes, y_true_masks, y_pred_boxes, y_pred_classes, alpha, gamma, delta,
                  with_grad=True)


class EfficientDetNetTrain(object):
  '''Only the head-loss part of the reference class: `_get_loss` with the reference's argument order.'''

  def __init__(self, alpha=0.25, gamma=1.5, **unused):
    self.alpha = alpha
    self.gamma = gamma

  def _get_loss(self, y_true_boxes, y_true_classes, y_true_masks, y_pred_boxes, y_pred_classes):
    return get_loss(y_true_boxes, y_true_classes, y_true_masks, y_pred_boxes, y_pred_classes, self.alpha, self.gamma)
