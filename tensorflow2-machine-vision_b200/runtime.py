"""CUDA-graph capture for launch-bound steps (B200-first: streams and graphs, no tracing compiler).

A step of this path is 8-10 small launches; issued one by one from Python the GPU idles between them.
``capture(fn)`` records them once into a CUDA graph on a side stream (the C-ABI calls enqueue on torch's current
stream and never allocate or synchronise, so they are capturable) and returns a callable that replays the graph.
Inputs must be static tensors that are updated in place between replays.

The data-parallel loss exchange is part of the captured step too: ``PeerExchange`` sets up the NVLink peer mailboxes
the ``*_dp`` entry points use inside their finalize kernels (no collective launch at all), ``NcclExchange`` a private
NCCL communicator for ``b200_allreduce_loss`` (ncclAllReduce issued by the library on the capture stream).
torch.distributed is used for the one-time rendezvous only (handle / unique-id exchange), never on the step path.
"""
import ctypes

import torch


class CapturedStep(object):
    def __init__(self, fn, warmup=2):
        self.fn = fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):  # allocator + lazy module loading must settle before capture
                self.out = fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: other threads of the process (e.g. NCCL's watchdog) may keep calling the CUDA runtime
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.out = fn()

    def __call__(self):
        self.graph.replay()
        return self.out


def capture(fn, warmup=2):
    return CapturedStep(fn, warmup=warmup)


def _dist_info(group=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


class PeerExchange(object):
    """NVLink peer mailboxes of one data-parallel group (one process per GPU): create, exchange the CUDA-IPC handles
    through torch.distributed once, map the peers.  ``args()`` = (rank, world, mailboxes) for the *_dp entry points."""

    def __init__(self, group=None):
        import torch.distributed as dist
        from . import _lib
        self.lib = _lib.load()
        self.rank, self.world = _dist_info(group)
        self.own = ctypes.c_void_p()
        self.mapped = []
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(self.lib.b200_peer_mailbox_create(ctypes.byref(self.own), handle), "PeerExchange")
        ptrs = [None] * 8
        if self.world > 1:
            if self.world > 8:
                raise ValueError("PeerExchange supports at most 8 ranks (one NVSwitch domain)")
            blobs = [None] * self.world
            dist.all_gather_object(blobs, bytes(bytearray(handle)), group=group)
            for r in range(self.world):
                if r == self.rank:
                    ptrs[r] = self.own.value
                    continue
                m = ctypes.c_void_p()
                h = (ctypes.c_ubyte * 64)(*blobs[r])
                _lib.check(self.lib.b200_peer_mailbox_open(h, ctypes.byref(m)), "PeerExchange (CUDA IPC open of rank %d)" % r)
                self.mapped.append(m)
                ptrs[r] = m.value
            dist.barrier(group=group)   # every mailbox is mapped everywhere before the first exchange
        else:
            ptrs[0] = self.own.value
        self.mailboxes = (ctypes.c_void_p * 8)(*ptrs)

    def args(self):
        return self.rank, self.world, self.mailboxes

    def status(self):
        """(exchanges completed, timeouts) of this rank's mailbox; synchronises."""
        from . import _lib
        e, n = ctypes.c_ulonglong(0), ctypes.c_uint(0)
        _lib.check(self.lib.b200_peer_mailbox_status(self.own, ctypes.byref(e), ctypes.byref(n)), "PeerExchange.status")
        return int(e.value), int(n.value)

    def allreduce_(self, t):
        """In-place sum over the ranks of a small fp32 / fp64 CUDA tensor (<= 32 values): the stand-alone kernel."""
        from . import _lib, _tensors as T
        fn = self.lib.b200_allreduce_sums_peer if t.dtype == torch.float64 else self.lib.b200_allreduce_loss_peer
        if t.dtype not in (torch.float32, torch.float64) or not t.is_cuda or not t.is_contiguous():
            raise ValueError("allreduce_ takes a contiguous fp32/fp64 CUDA tensor")
        _lib.check(fn(t.data_ptr(), t.numel(), self.rank, self.world, self.mailboxes, T.stream_ptr()), "PeerExchange.allreduce_")
        return t

    def collect_yolo(self, parts_out, loss_out):
        """Second half of a deferred GetLoss exchange (_loss_call(..., defer_collect=True)): waits for the peers' terms,
        writes the global parts (3,4) / loss.  Enqueue it on a stream ordered behind the publishing call; the publish of
        step f must in turn be ordered behind the collect of step f-2 (four slot sets)."""
        from . import _lib, _tensors as T
        _lib.check(self.lib.b200_yolo_loss_collect_peer(T.ptr(parts_out), T.ptr(loss_out), self.rank, self.world, self.mailboxes,
                                                        T.stream_ptr()), "PeerExchange.collect_yolo")

    def close(self):
        for m in self.mapped:
            self.lib.b200_peer_mailbox_close(m)
        self.mapped = []
        if self.own:
            self.lib.b200_peer_mailbox_destroy(self.own)
            self.own = ctypes.c_void_p()


class NcclExchange(object):
    """A private NCCL communicator for b200_allreduce_loss / b200_allreduce_sums (rank 0 makes the unique id, broadcast
    once through torch.distributed)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        from . import _lib
        self.lib = _lib.load()
        self.rank, self.world = _dist_info(group)
        idb = (ctypes.c_ubyte * 128)()
        if self.rank == 0:
            _lib.check(self.lib.b200_nccl_unique_id(idb), "NcclExchange")
        blob = [bytes(bytearray(idb))]
        if self.world > 1:
            dist.broadcast_object_list(blob, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        idb = (ctypes.c_ubyte * 128)(*blob[0])
        self.comm = ctypes.c_void_p()
        _lib.check(self.lib.b200_nccl_comm_init(ctypes.byref(self.comm), self.world, self.rank, idb), "NcclExchange")

    def allreduce_(self, t):
        from . import _lib, _tensors as T
        fn = self.lib.b200_allreduce_sums if t.dtype == torch.float64 else self.lib.b200_allreduce_loss
        if t.dtype not in (torch.float32, torch.float64) or not t.is_cuda or not t.is_contiguous():
            raise ValueError("allreduce_ takes a contiguous fp32/fp64 CUDA tensor")
        _lib.check(fn(self.comm, t.data_ptr(), t.numel(), T.stream_ptr()), "NcclExchange.allreduce_")
        return t

    def close(self):
        if self.comm:
            self.lib.b200_nccl_comm_destroy(self.comm)
            self.comm = ctypes.c_void_p()
