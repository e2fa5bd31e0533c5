"""CUDA-graph replay of one Os2dHead call for a fixed feature-map shape.

The six kernels of a head call are launch-bound for small workloads (512 px input, a handful of classes: ~0.2 ms of GPU
work behind ~6 launches, workspace allocations and tensor-map encodes).  ``GraphedHead`` captures one call - kernels,
workspace and outputs live in the graph's private memory pool - and replays it.  Opt-in, because it changes the ownership
convention of the reference API: the returned tensors are the graph's static buffers and are overwritten by the next call
(``Os2dHead.forward`` itself always returns fresh tensors, evaluate.py:351-357 keeps them across calls).
"""
import torch


class GraphedHead:
    def __init__(self, head, feature_maps_like, warmup=3):
        """head: os2d_b200.head.Os2dHead; feature_maps_like: a CUDA tensor [B,D,H,W] fixing shape, dtype and device."""
        if feature_maps_like.device.type != "cuda":
            raise RuntimeError("os2d_b200 requires CUDA tensors (no CPU path)")
        self.head = head
        self.static_in = torch.empty_like(feature_maps_like, dtype=torch.float32).contiguous()
        self.static_in.copy_(feature_maps_like)
        side = torch.cuda.Stream(device=self.static_in.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(1, warmup)):      # one-time work (function attributes, driver entry points) outside the capture
                head(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = head(self.static_in)

    def __call__(self, feature_maps):
        """Same return tuple as Os2dHead.forward; the tensors are reused by the next call."""
        assert feature_maps.shape == self.static_in.shape, "GraphedHead was captured for shape {}".format(tuple(self.static_in.shape))
        self.static_in.copy_(feature_maps, non_blocking=True)
        self.graph.replay()
        return self.static_out
