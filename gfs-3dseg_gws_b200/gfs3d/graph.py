"""CUDA-graph capture of the eval forward: ~70 kernel launches of one step collapse into one graph launch, so the step is
immune to host-side launch latency (no tracing compiler involved: the captured work is exactly the hand-written kernels
plus the few tensor ops of the eager path)."""
import torch


class GraphedEval:
    """fn(x) -> tensor(s), captured once for a fixed input shape; call with any tensor of that shape"""

    def __init__(self, fn, example: torch.Tensor, warmup: int = 3):
        self.static_in = example.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):          # warm-up off the default stream: one-time attribute / cache set-up happens here
                fn(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        from . import ops
        n0 = ops.LAUNCHES
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(self.static_in)
        self.kernels_per_replay = ops.LAUNCHES - n0      # hand-written kernels captured per replay

    def __call__(self, x: torch.Tensor, non_blocking: bool = True):
        self.static_in.copy_(x, non_blocking=non_blocking)
        self.graph.replay()
        return self.static_out
