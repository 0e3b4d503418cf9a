"""Device-memory plumbing for the shims (torch is used for allocation, streams and DLPack only)."""
import numpy as np
import torch


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("tensorflow2-machine-vision_b200 needs a CUDA device (B200); there is no CPU fallback")


def to_cuda(x, dtype=torch.float32):
    """Anything tensor-like -> contiguous CUDA tensor of `dtype` (zero-copy for CUDA/DLPack inputs)."""
    require_cuda()
    if isinstance(x, torch.Tensor):
        t = x
    elif hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray):
        t = torch.from_dlpack(x)  # e.g. tf.experimental.dlpack / cupy
    else:
        a = np.ascontiguousarray(np.asarray(x))
        t = torch.from_numpy(a)
        if t.numel() > (1 << 16):
            t = t.pin_memory()
    if t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_cuda:
        t = t.cuda(non_blocking=True)
    return t.contiguous()


def is_pinned_host_f32(x):
    """A contiguous fp32 torch tensor in page-locked host memory: device kernels can read it in place (UVA)."""
    return (isinstance(x, torch.Tensor) and not x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
            and x.is_pinned() and torch.cuda.is_available())


def ptr(t):
    return 0 if t is None else t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def host_floats(x, n=None):
    """Small host-side float32 parameter (anchors, image size) as a flat numpy array."""
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float32)).reshape(-1)
    if n is not None and a.size != n:
        raise ValueError("expected %d values, got %d" % (n, a.size))
    return a
