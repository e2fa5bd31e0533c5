"""Multi-GPU: shard the class (query) axis over the ranks of one node, one process per GPU.

In eval mode every class's correlation / TransformNet / resample / box regression is independent
(reference: per-class loop in os2d/engine/evaluate.py:323-327, per-class NMS box_coder.py:483-528), so the
class axis partitions with no data-path exchange; the only collective is one all-gather of the per-class
outputs before NMS (BASELINE.json north_star).  Each rank writes its class block [B, C_local, 13, N]
(score 1 + loc 4 + corners 8 planes) directly into its slice of the gather buffer; the all-gather runs in
place on that buffer (NCCL over NVLink on GPUs, gloo in the CPU unit tests).
"""
import torch
import torch.distributed as dist

OUT_PLANES = 13   # score(1) + loc(4) + corners(8)


def shard_bounds(num_classes, world_size, rank):
    """Contiguous block [lo, hi) of classes owned by ``rank``; blocks of ceil(C/world) classes (the last ranks may
    own fewer or none)."""
    per = -(-num_classes // world_size)
    lo = min(rank * per, num_classes)
    return lo, min(lo + per, num_classes)


def padded_block(num_classes, world_size):
    return -(-num_classes // world_size)


def allocate_gather_buffer(B, num_classes, N, world_size, device):
    """[world, B, per, 13, N] fp32: rank r's block is buffer[r]; classes beyond num_classes are padding."""
    per = padded_block(num_classes, world_size)
    return torch.zeros(world_size, B, per, OUT_PLANES, N, dtype=torch.float32, device=device)


def local_views(buffer, rank):
    """(score [B,per,1,N], loc [B,per,4,N], corners [B,per,8,N]) views of this rank's block."""
    blk = buffer[rank]
    return blk[:, :, 0:1], blk[:, :, 1:5], blk[:, :, 5:13]


def all_gather_outputs(buffer, group=None, async_op=False):
    """In-place all-gather of every rank's block of ``buffer`` ([world, ...]).  With ``async_op`` the NCCL work handle is
    returned so that the collective of image i overlaps the kernels of image i+1 (wait before reading / reusing)."""
    world = dist.get_world_size(group)
    if world == 1:
        return None if async_op else buffer
    rank = dist.get_rank(group)
    src = buffer[rank].reshape(-1)
    if buffer.device.type == "cpu":
        src = src.clone()          # gloo does not support the in-place form
    work = dist.all_gather_into_tensor(buffer.view(-1), src, group=group, async_op=async_op)
    return work if async_op else buffer


def unpack_gathered(buffer, num_classes):
    """[world,B,per,13,N] -> (loc [B,C,4,N], score [B,C,N], corners [B,C,8,N]) in global class order."""
    world, B, per, _, N = buffer.shape
    full = buffer.permute(1, 0, 2, 3, 4).reshape(B, world * per, OUT_PLANES, N)[:, :num_classes]
    return full[:, :, 1:5], full[:, :, 0], full[:, :, 5:13]


class SymmetricGatherBuffer:
    """[world, B, per, 13, N] gather buffer in symmetric memory (torch.distributed._symmetric_memory): every rank maps the
    buffers of all ranks, so K3 can store its outputs into all of them (csrc/resample_p2p.cu) and a device-side barrier
    replaces the all-gather.  NOT yet validated on a multi-GPU box; ``ClassShardedHead(fused_gather=True)`` opts in."""

    def __init__(self, B, num_classes, N, world_size, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.per = padded_block(num_classes, world_size)
        self.shape = (world_size, B, self.per, OUT_PLANES, N)
        self.buffer = symm_mem.empty(*self.shape, dtype=torch.float32, device=device)
        self.buffer.zero_()
        self.handle = symm_mem.rendezvous(self.buffer, group if group is not None else dist.group.WORLD)
        self.ptrs = torch.tensor(list(self.handle.buffer_ptrs), dtype=torch.int64, device=device)
        self.slice_elems = B * self.per * OUT_PLANES * N

    def out_peers(self, rank):
        return self.ptrs, rank * self.slice_elems, self.per

    def barrier(self):
        """All ranks have finished their K3 stores (stream-ordered, system-scope release / acquire)."""
        self.handle.barrier()


class ClassShardedHead:
    """Runs an ``Os2dHead`` built from this rank's class block and all-gathers the per-class outputs.

    ``head_factory(class_maps_block)`` creates the local head (normally
    ``os2d_head_creator.create_os2d_head``); it is only called when the block is not empty.
    ``fused_gather=True`` (experimental): the head's K3 writes into every rank's symmetric gather buffer itself and a
    device-side barrier replaces the NCCL all-gather.
    """

    def __init__(self, class_feature_maps, head_factory, group=None, fused_gather=False):
        self.fused_gather = bool(fused_gather)
        self._symm = {}
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.num_classes = len(class_feature_maps)
        lo, hi = shard_bounds(self.num_classes, self.world, self.rank)
        self.lo, self.hi = lo, hi
        self.head = head_factory(class_feature_maps[lo:hi]) if hi > lo else None

    def forward(self, feature_maps):
        B, _, H, W = feature_maps.shape
        N = H * W
        if self.fused_gather and self.world > 1:
            return self._forward_fused(feature_maps, B, H, W, N)
        buf = allocate_gather_buffer(B, self.num_classes, N, self.world, feature_maps.device)
        if self.head is not None:
            s_v, l_v, c_v = local_views(buf, self.rank)
            n = self.hi - self.lo
            if getattr(self.head, "supports_out_views", False):
                # the resample kernel writes straight into this rank's slice of the gather buffer (no staging copy)
                self.head(feature_maps, out_views=(s_v[:, :n], l_v[:, :n], c_v[:, :n]))
            else:
                loc, score, _, corners = self.head(feature_maps)
                s_v[:, :n].copy_(score.reshape(B, n, 1, N))
                l_v[:, :n].copy_(loc.reshape(B, n, 4, N))
                c_v[:, :n].copy_(corners.reshape(B, n, 8, N))
        if self.world > 1:
            all_gather_outputs(buf, self.group)
        loc, score, corners = unpack_gathered(buf, self.num_classes)
        return (loc.reshape(B, self.num_classes, 4, H, W), score.reshape(B, self.num_classes, 1, H, W),
                corners.reshape(B, self.num_classes, 8, H, W))

    def _forward_fused(self, feature_maps, B, H, W, N):
        key = (B, N)
        sg = self._symm.get(key)
        if sg is None:                                   # collective allocation + rendezvous, once per shape
            sg = self._symm[key] = SymmetricGatherBuffer(B, self.num_classes, N, self.world, feature_maps.device, self.group)
        sg.barrier()                                     # every rank is done reading the previous contents
        if self.head is not None:
            self.head(feature_maps, out_peers=sg.out_peers(self.rank))
        sg.barrier()                                     # every rank's stores have landed everywhere
        loc, score, corners = unpack_gathered(sg.buffer, self.num_classes)
        # the buffer is reused by the next call: hand out copies (the reference API returns fresh tensors)
        return (loc.reshape(B, self.num_classes, 4, H, W).clone(), score.reshape(B, self.num_classes, 1, H, W).clone(),
                corners.reshape(B, self.num_classes, 8, H, W).clone())

    __call__ = forward
