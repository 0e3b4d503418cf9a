"""Mirror of the one function of the reference's ImageHelper that sits on the serving path:
opencvProportionalResize (utils/image_helper.py:293-325) — proportional INTER_AREA resize + constant border — on the
device (csrc/letterbox.cu, OpenCV's 8-bit arithmetic bit for bit).  The augmentation options of the reference
(random background colour, BORDER_REPLICATE) are not part of the serving path and are refused."""
import ctypes

import numpy as np
import torch

from ... import _lib, _tensors as T

BORDER_CONSTANT = 0  # cv2.BORDER_CONSTANT


def opencvGetImageSize(opencv_img):
  '''(width, height) of an HxWxC image (image_helper.py:82-86)'''
  return int(opencv_img.shape[1]), int(opencv_img.shape[0])


def _image_on_device(opencv_img):
  T.require_cuda()
  if isinstance(opencv_img, torch.Tensor):
    t = opencv_img
  else:
    t = torch.from_numpy(np.ascontiguousarray(opencv_img))
  if t.dtype != torch.uint8 or t.dim() != 3:
    raise ValueError('expected an HxWx3 uint8 image, got %s %s' % (tuple(t.shape), t.dtype))
  if not t.is_cuda and not (t.is_pinned() and t.is_contiguous()):
    t = t.cuda()  # pageable host memory: one copy; pinned host memory is read in place by the kernel
  return t.contiguous()


def letterbox(opencv_img, size, bg_color, want_u8, want_f32):
  '''One launch of b200_letterbox_image -> (uint8 image or None, float32 RGB/255 image or None, padding, (rw, rh))'''
  lib = _lib.load()
  img = _image_on_device(opencv_img)
  dev = torch.device('cuda', torch.cuda.current_device()) if not img.is_cuda else img.device
  new_width, new_height = int(size[0]), int(size[1])
  out_u8 = torch.empty((new_height, new_width, 3), dtype=torch.uint8, device=dev) if want_u8 else None
  out_f32 = torch.empty((new_height, new_width, 3), dtype=torch.float32, device=dev) if want_f32 else None
  bg = (ctypes.c_uint8 * 3)(*[int(min(max(round(float(v)), 0), 255)) for v in list(bg_color)[:3]])
  padding = (ctypes.c_int32 * 4)()
  resized = (ctypes.c_int32 * 2)()
  _lib.check(lib.b200_letterbox_image(T.ptr(img), int(img.shape[0]), int(img.shape[1]), int(img.shape[2]), new_width, new_height,
                                      bg, T.ptr(out_u8), T.ptr(out_f32), padding, resized, T.stream_ptr()), 'opencvProportionalResize')
  return out_u8, out_f32, tuple(int(v) for v in padding), (int(resized[0]), int(resized[1]))


def opencvProportionalResize(opencv_img, size, points=None, bg_color=(128, 128, 128), bg_mode=BORDER_CONSTANT):
  '''
  Args:
    opencv_img: HxWx3 uint8 (numpy, or a torch tensor on the device / in pinned host memory); size: (new_width, new_height)
    points: optional [[x, y], ...] on the original image
  Returns:
    result_img (new_height,new_width,3) uint8 on the device, result_points float32 (n,2), padding (top,bottom,left,right)
  '''
  if bg_color is None or bg_mode != BORDER_CONSTANT:
    raise NotImplementedError('opencvProportionalResize: only a given colour with BORDER_CONSTANT (the serving path) is built')
  width, height = opencvGetImageSize(opencv_img)
  result_img, _, padding, (resize_width, resize_height) = letterbox(opencv_img, size, bg_color, True, False)
  result_points = []
  if points is not None:
    for p in points:  # image_helper.py:319-323
      result_points.append([p[0] * resize_width / width + padding[2], p[1] * resize_height / height + padding[0]])
  return result_img, np.float32(result_points), padding
