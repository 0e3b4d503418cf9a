"""CUDA-graph capture for launch-bound steps (B200-first: streams and graphs, no tracing compiler).

A step of this path is 8-10 small launches; issued one by one from Python the GPU idles between them.
``capture(fn)`` records them once into a CUDA graph on a side stream (the C-ABI calls enqueue on torch's current
stream and never allocate or synchronise, so they are capturable, NCCL all-reduce included) and returns a callable
that replays the graph.  Inputs must be static tensors that are updated in place between replays.
"""
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
