"""Multi-GPU: shard the class (query) axis over the ranks of one node, one process per GPU.

In eval mode every class's correlation / TransformNet / resample / box regression is independent
(reference: per-class loop in os2d/engine/evaluate.py:323-327, per-class NMS box_coder.py:483-528), so the
class axis partitions with no data-path exchange; the only collective is one all-gather of the per-class
outputs before NMS (BASELINE.json north_star).  Each rank's K3 writes its class block [B, C_local, 13, N]
(score 1 + loc 4 + corners 8 planes) directly into its slice of a persistent, double-buffered gather buffer.

Three ways to run that all-gather (``ClassShardedHead(gather=...)``):

  "copy_engine" (default on GPUs)  the buffers live in symmetric memory (every rank maps every rank's buffer over NVLink);
                 each rank PUSHES its slice into the peers' buffers with cudaMemcpyAsync on a side stream - the DMA copy
                 engines move the bytes, ZERO SMs are taken from the persistent tcgen05 kernels of the next image (an NCCL
                 all-gather runs SM-resident CTAs that delay the 148-CTA correlation kernel: 0.25 -> 0.45 ms at 8 GPUs in
                 round 1).  Two device-side barriers (symmetric-memory signal pads, stream-ordered) bracket the pushes.
  "fused"        K3 itself stores its outputs into every rank's buffer (csrc/resample.cu PeerSink): compute and
                 collective in one kernel, a barrier before (slot free everywhere) and after (stores landed).
  "nccl"         in-place ncclAllGather, asynchronous (also the gloo path of the CPU tests).

``submit()`` is asynchronous and pipelined: the gather of image i overlaps the kernels of image i+1; ``forward()`` =
``submit().wait()`` + fresh output tensors (the reference API contract).

Also here: ``ShardedUpload`` (every rank uploads 1/G of a replicated host tensor over its own PCIe link and the parts are
exchanged over NVLink) and ``ClassShardedDetector`` (decode + NMS on each rank's own labels, survivors gathered).
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
    """In-place all-gather of every rank's block of ``buffer`` ([world, ...]).  With ``async_op`` the work handle is
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
    """[world,B,per,13,N] -> (loc [B,C,4,N], score [B,C,N], corners [B,C,8,N]) in global class order (views)."""
    world, B, per, _, N = buffer.shape
    full = buffer.permute(1, 0, 2, 3, 4).reshape(B, world * per, OUT_PLANES, N)[:, :num_classes]
    return full[:, :, 1:5], full[:, :, 0], full[:, :, 5:13]


class SymmetricSlots:
    """``depth`` slots of a [world, *part_shape] tensor in symmetric memory (torch.distributed._symmetric_memory): slot k of
    every rank is mapped by every rank, part r of a slot is produced by rank r.  ``exchange(k, ...)`` pushes this rank's part
    of slot k into all peers with the copy engines on ``stream`` between two device-side barriers."""

    def __init__(self, depth, world, rank, part_shape, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.depth, self.world, self.rank = depth, world, rank
        self.shape = (depth, world) + tuple(part_shape)
        self.buffer = symm_mem.empty(*self.shape, dtype=dtype, device=device)
        self.buffer.zero_()
        self.handle = symm_mem.rendezvous(self.buffer, group if group is not None else dist.group.WORLD)
        self.peers = [self.handle.get_buffer(r, self.shape, dtype) if r != rank else self.buffer for r in range(world)]
        self.slot_elems = self.buffer[0].numel()
        self.part_elems = self.buffer[0, 0].numel()
        self.stream = torch.cuda.Stream(device=device)
        torch.cuda.synchronize(device)
        self.handle.barrier(channel=0)           # every rank has zeroed its buffer before anybody pushes into it
        torch.cuda.synchronize(device)

    def peer_slot_pointers(self, k):
        """int64 CUDA tensor of the base addresses of slot k in every rank's buffer (for the fused K3 epilogue)."""
        esz = self.buffer.element_size()
        ptrs = [int(p) + k * self.slot_elems * esz for p in self.handle.buffer_ptrs]
        return torch.tensor(ptrs, dtype=torch.int64, device=self.buffer.device)

    def barrier(self, k):
        """Stream-ordered (current stream) barrier across the ranks, system-scope release / acquire."""
        self.handle.barrier(channel=k % 8)

    def exchange(self, k, after_event, fill=None):
        """On the side stream: wait for ``after_event`` (this rank's producer of slot k, and everything before it on the
        producer stream - including the consumer of the slot's previous contents), barrier (slot k is free on every rank),
        optional ``fill()`` (e.g. an H2D copy of this rank's part), push this rank's part to every peer, barrier (all
        parts have landed everywhere).  Returns the event that marks slot k complete on this rank."""
        with torch.cuda.stream(self.stream):
            for ev in (after_event if isinstance(after_event, (list, tuple)) else [after_event]):
                if ev is not None:
                    self.stream.wait_event(ev)
            self.barrier(k)
            if fill is not None:
                fill()
            src = self.buffer[k, self.rank]
            for d in range(1, self.world):
                r = (self.rank + d) % self.world           # staggered targets: no two ranks push to the same peer at once
                self.peers[r][k, self.rank].copy_(src, non_blocking=True)
            self.barrier(k)
            done = torch.cuda.Event()
            done.record(self.stream)
        return done


class GatherHandle:
    """Result of ``ClassShardedHead.submit``: ``wait()`` makes the current stream wait for the gather and returns views
    (loc [B,C,4,H,W], score [B,C,1,H,W], corners [B,C,8,H,W]) of the slot, valid until ``depth`` further submits."""

    def __init__(self, owner, slot, pending, B, H, W):
        self.owner, self.slot, self.pending, self.B, self.H, self.W = owner, slot, pending, B, H, W
        self.local_event = None     # set when this rank's own block is produced on a side stream (async resample)

    def wait(self):
        if isinstance(self.pending, str):           # exchange still deferred (see ClassShardedHead.submit): start it now
            self.owner._flush_deferred()
        p = self.pending
        if p is not None:
            if isinstance(p, torch.cuda.Event):
                torch.cuda.current_stream().wait_event(p)
            else:
                p.wait()
            self.pending = None
        C = self.owner.num_classes
        loc, score, corners = unpack_gathered(self.slot, C)
        B, H, W = self.B, self.H, self.W
        return loc.reshape(B, C, 4, H, W), score.reshape(B, C, 1, H, W), corners.reshape(B, C, 8, H, W)

    def local_views(self):
        """This rank's own block (score [B,per,1,N], loc [B,per,4,N], corners [B,per,8,N]): complete as soon as the head
        has run on the submitting stream, independent of the gather."""
        return local_views(self.slot, self.owner.rank)


class ClassShardedHead:
    """Runs an ``Os2dHead`` built from this rank's class block and gathers the per-class outputs of all ranks.

    ``head_factory(class_maps_block)`` creates the local head (normally ``os2d_head_creator.create_os2d_head``); it is only
    called when the block is not empty.  ``gather``: "copy_engine" | "fused" | "nccl" (see the module docstring); None picks
    "copy_engine" on CUDA with more than one rank, else "nccl" (gloo on CPU).  ``depth``: gather buffers in flight.
    Results must be consumed on (or their consumer joined to) the stream that issues the next ``submit`` to the same slot.
    """

    def __init__(self, class_feature_maps, head_factory, group=None, gather=None, depth=2, fused_gather=False):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.num_classes = len(class_feature_maps)
        lo, hi = shard_bounds(self.num_classes, self.world, self.rank)
        self.lo, self.hi = lo, hi
        self.head = head_factory(class_feature_maps[lo:hi]) if hi > lo else None
        if fused_gather:
            gather = "fused"
        self.gather = gather
        self.depth = depth
        self._rings = {}
        self._step = 0
        # copy-engine mode, optional: enqueue the pushes of image i only once K1 of image i+1 has been launched (and wait for
        # it), so that the inbound NVLink writes overlap conv1 instead of the correlation kernel.  Measured at 8 GPUs: K1 is
        # back at 0.25 ms (from 0.32), but the two all-rank barriers now also wait for every rank's K1 and the step gets
        # LONGER (2.11 -> 2.39 ms, profiles/r02_bench_n8_*_run2_deferred_exchange.json) - so it is off by default.
        self.defer_exchange = False
        self._deferred = None
        # copy-engine mode: launch K3 through Os2dHead.submit (side stream), so it overlaps the next image's tensor kernels
        self.async_resample = False

    # ---- buffers: allocated once per (B, N) shape (collective for the symmetric modes), reused by every call ----
    def _ring(self, B, N, device):
        key = (B, N)
        ring = self._rings.get(key)
        if ring is None:
            mode = self.gather
            if mode is None:
                mode = "copy_engine" if (device.type == "cuda" and self.world > 1) else "nccl"
            if self.world == 1 or device.type != "cuda":
                mode = "nccl"
            per = padded_block(self.num_classes, self.world)
            ring = {"mode": mode, "pending": [None] * self.depth, "ptrs": None}
            if mode in ("copy_engine", "fused"):
                # symmetric memory needs NVLink peer mapping between all ranks; where the platform refuses it, every rank
                # agrees (one all-reduce) to run this ring over the NCCL transport instead - loudly, never silently
                slots, err = None, None
                try:
                    slots = SymmetricSlots(self.depth, self.world, self.rank, (B, per, OUT_PLANES, N), torch.float32,
                                           device, self.group)
                except Exception as e:   # noqa: BLE001
                    err = e
                ok = torch.tensor([0 if slots is None else 1], dtype=torch.int32, device=device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
                if int(ok.item()) == 0:
                    import warnings
                    warnings.warn("os2d_b200.dist: symmetric memory is not available on this platform ({}); the '{}' gather "
                                  "runs over NCCL instead".format(err, mode))
                    mode = ring["mode"] = "nccl"
                    slots = None
            if mode in ("copy_engine", "fused"):
                ring["slots"] = slots
                ring["buffer"] = ring["slots"].buffer
                if mode == "fused":
                    ring["ptrs"] = [ring["slots"].peer_slot_pointers(k) for k in range(self.depth)]
            else:
                ring["buffer"] = torch.zeros(self.depth, self.world, B, per, OUT_PLANES, N, dtype=torch.float32, device=device)
            self._rings[key] = ring
        return ring

    def submit(self, feature_maps):
        """Enqueue the local head and the (asynchronous) gather for one image batch; returns a ``GatherHandle``."""
        B, _, H, W = feature_maps.shape
        N = H * W
        ring = self._ring(B, N, feature_maps.device)
        k = self._step % self.depth
        self._step += 1
        slot = ring["buffer"][k]
        if self._deferred is not None and (self._deferred["ring"] is not ring or self._deferred["k"] == k):
            self._flush_deferred()
        prev = ring["pending"][k]
        if prev is not None:                       # the previous gather into this slot must be over before it is rewritten
            if isinstance(prev, torch.cuda.Event):
                torch.cuda.current_stream().wait_event(prev)
            else:
                prev.wait()
            ring["pending"][k] = None
        n = self.hi - self.lo
        mode = ring["mode"]
        if mode == "fused":
            slots = ring["slots"]
            if self.head is not None:
                per = slot.shape[2]
                self.head(feature_maps, out_peers=(ring["ptrs"][k], self.rank * slots.part_elems, per),
                          _before_resample=lambda: slots.barrier(k))     # slot k is free on every rank
            else:
                slots.barrier(k)
            slots.barrier(k)                                              # every rank's stores have landed everywhere
            return GatherHandle(self, slot, None, B, H, W)
        ev_k3 = None

        def after_corr():
            if self._deferred is not None:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                self._flush_deferred(ev)

        if self.head is not None:
            s_v, l_v, c_v = local_views(slot, self.rank)
            if getattr(self.head, "supports_out_views", False):
                # the resample kernel writes straight into this rank's slice of the gather buffer (no staging copy)
                views = (s_v[:, :n], l_v[:, :n], c_v[:, :n])
                if self.async_resample and mode == "copy_engine" and hasattr(self.head, "submit"):
                    _, ev_k3 = self.head.submit(feature_maps, out_views=views, _after_corr=after_corr)
                else:
                    self.head(feature_maps, out_views=views, _after_corr=after_corr)
            else:
                loc, score, _, corners = self.head(feature_maps)
                s_v[:, :n].copy_(score.reshape(B, n, 1, N))
                l_v[:, :n].copy_(loc.reshape(B, n, 4, N))
                c_v[:, :n].copy_(corners.reshape(B, n, 8, N))
        after_corr()                               # no-op when the head already triggered it
        pending = None
        if self.world > 1:
            if mode == "copy_engine":
                ev = ev_k3                         # this rank's block is complete when K3 is (side stream) ...
                if ev is None:
                    ev = torch.cuda.Event()        # ... or when the submitting stream gets here
                    ev.record(torch.cuda.current_stream())
                if self.defer_exchange:
                    handle = GatherHandle(self, slot, "deferred", B, H, W)
                    handle.local_event = ev_k3
                    self._deferred = {"ring": ring, "k": k, "ev_head": ev, "handle": handle}
                    return handle
                pending = ring["slots"].exchange(k, ev)
            else:
                pending = all_gather_outputs(slot, self.group, async_op=True)
            ring["pending"][k] = pending
        handle = GatherHandle(self, slot, pending, B, H, W)
        handle.local_event = ev_k3
        return handle

    def _flush_deferred(self, extra_event=None):
        d = self._deferred
        if d is None:
            return
        self._deferred = None
        ev = d["ring"]["slots"].exchange(d["k"], [d["ev_head"], extra_event])
        d["ring"]["pending"][d["k"]] = ev
        d["handle"].pending = ev

    def transport(self):
        """The gather transports actually in use, per (B, N) ring: 'copy_engine' | 'fused' | 'nccl'."""
        return sorted(set(r["mode"] for r in self._rings.values()))

    def drain(self):
        """Wait (stream-level) for every gather still in flight."""
        self._flush_deferred()
        for ring in self._rings.values():
            for k, p in enumerate(ring["pending"]):
                if p is not None:
                    if isinstance(p, torch.cuda.Event):
                        torch.cuda.current_stream().wait_event(p)
                    else:
                        p.wait()
                    ring["pending"][k] = None

    def forward(self, feature_maps):
        """Synchronous form with the reference's contract (fresh output tensors): (loc, score, corners) over ALL classes."""
        loc, score, corners = self.submit(feature_maps).wait()
        return loc.clone(), score.clone(), corners.clone()

    __call__ = forward


class ShardedUpload:
    """Host -> device upload of a tensor that every rank holds on the host (the replicated image feature map): rank r copies
    only part r (1/G of the bytes) over its own PCIe link into a symmetric buffer and the parts are exchanged over NVLink by
    the copy engines.  Double buffered; ``upload`` returns (device tensor, event to wait for)."""

    def __init__(self, shape, dtype, device, group=None, depth=2):
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.shape, self.dtype, self.depth = tuple(shape), dtype, depth
        numel = 1
        for s in shape:
            numel *= s
        self.numel = numel
        self.part = -(-numel // self.world)
        self.part = -(-self.part // 64) * 64
        self.slots = None
        self.full_copy = self.world == 1               # every rank copies the whole tensor itself
        if self.world > 1:
            err = None
            try:
                self.slots = SymmetricSlots(depth, self.world, self.rank, (self.part,), dtype, device, group)
            except Exception as e:   # noqa: BLE001
                err = e
            ok = torch.tensor([0 if self.slots is None else 1], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:                     # agreed by all ranks: no peer mapping here, upload everything locally
                import warnings
                warnings.warn("os2d_b200.dist.ShardedUpload: symmetric memory is not available ({}); every rank uploads the "
                              "whole tensor".format(err))
                self.slots, self.full_copy = None, True
        if self.slots is not None:
            self.flat = [self.slots.buffer[k].view(-1) for k in range(depth)]
        else:
            self.flat = [torch.empty(max(self.part, numel), dtype=dtype, device=device) for _ in range(depth)]
            self.stream = torch.cuda.Stream(device=device)
        self._step = 0

    def _own_range(self):
        if self.full_copy:
            return 0, self.numel
        lo = min(self.rank * self.part, self.numel)
        return lo, min(lo + self.part, self.numel)

    def bytes_per_rank(self):
        lo, hi = self._own_range()
        return (hi - lo) * torch.empty(0, dtype=self.dtype).element_size()

    def upload(self, host_pinned, after_event=None):
        """``after_event``: the consumer of the slot's previous contents (two uploads ago) on this rank is done."""
        k = self._step % self.depth
        self._step += 1
        src = host_pinned.view(-1)
        lo, hi = self._own_range()
        out = self.flat[k][:self.numel].view(self.shape)
        if self.slots is None:
            with torch.cuda.stream(self.stream):
                if after_event is not None:
                    self.stream.wait_event(after_event)
                self.flat[k][lo:hi].copy_(src[lo:hi], non_blocking=True)
                done = torch.cuda.Event()
                done.record(self.stream)
            return out, done
        dst = self.slots.buffer[k, self.rank]

        def fill():
            if hi > lo:
                dst[:hi - lo].copy_(src[lo:hi], non_blocking=True)
        return out, self.slots.exchange(k, after_event, fill=fill)


class ClassShardedDetector:
    """Head + decode + NMS sharded by REAL LABEL (all class views of a label live on one rank, so the per-label NMS of
    box_coder.py:483-528 needs no exchange), survivors gathered: the only collective moves detections, not the
    [C,13,N] score maps.  ``forward`` returns the same BoxList on every rank, labels in rank-then-set order."""

    def __init__(self, class_feature_maps, class_ids, head_factory, box_coder, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.box_coder = box_coder
        labels = list(dict.fromkeys(int(c) for c in class_ids))        # first-occurrence order
        lo, hi = shard_bounds(len(labels), self.world, self.rank)
        mine = set(labels[lo:hi])
        self.view_index = [i for i, c in enumerate(class_ids) if int(c) in mine]
        self.class_ids = [int(class_ids[i]) for i in self.view_index]
        self.head = head_factory([class_feature_maps[i] for i in self.view_index]) if self.view_index else None

    def local_detections(self, feature_maps_pyramid, img_size_pyramid, nms_score_threshold=0.0, nms_iou_threshold=0.3,
                         inverse_box_transforms=None, _async=False):
        """Decode + NMS of this rank's labels for ONE image given as a pyramid of [1,D,H_l,W_l] feature maps."""
        if self.head is None:
            return None
        loc_p, cls_p, cor_p = [], [], []
        C = len(self.view_index)
        for fm in feature_maps_pyramid:
            loc, score, _, corners = self.head(fm)
            n = fm.shape[2] * fm.shape[3]
            loc_p.append(loc[0].view(C, 4, n))
            cls_p.append(score[0].view(C, n))
            cor_p.append(corners[0].view(C, 8, n))
        decode = self.box_coder.decode_pyramid_async if _async else self.box_coder.decode_pyramid
        return decode(loc_p, cls_p, img_size_pyramid, self.class_ids, nms_score_threshold=nms_score_threshold,
                      nms_iou_threshold=nms_iou_threshold, inverse_box_transforms=inverse_box_transforms,
                      transform_corners_pyramid=cor_p)

    def submit(self, feature_maps_pyramid, img_size_pyramid, **kw):
        """Enqueue the heads and the decode + NMS kernel of one image; ``result(handle)`` finishes it.  Submitting image i+1
        before asking for the result of image i keeps the GPU busy while the host waits for image i's detection count."""
        pending = self.local_detections(feature_maps_pyramid, img_size_pyramid, _async=True, **kw)
        return pending, feature_maps_pyramid[0].device, img_size_pyramid, kw

    def forward(self, feature_maps_pyramid, img_size_pyramid, **kw):
        return self.result(self.submit(feature_maps_pyramid, img_size_pyramid, **kw))

    def result(self, handle):
        from .structures import BoxList
        from .box_coder import _probe_transform_target
        pending, device, img_size_pyramid, kw = handle
        dets = pending.result() if pending is not None else None
        if dets is not None:
            rows = torch.cat([dets.bbox_xyxy, dets.get_field("scores")[:, None],
                              dets.get_field("labels").to(torch.int32).view(torch.float32)[:, None],     # bit cast
                              dets.get_field("default_boxes").bbox_xyxy, dets.get_field("transform_corners")], dim=1)
        else:
            rows = torch.zeros(0, 18, dtype=torch.float32, device=device)
        inv = kw.get("inverse_box_transforms")
        image_size = img_size_pyramid[0] if inv is None else _probe_transform_target(inv[0], img_size_pyramid[0])
        if self.world > 1:
            cnt = torch.tensor([rows.shape[0]], dtype=torch.int64, device=device)
            counts = torch.empty(self.world, dtype=torch.int64, device=device)
            dist.all_gather_into_tensor(counts, cnt, group=self.group)
            counts = counts.tolist()
            cap = max(max(counts), 1)
            pad = torch.zeros(cap, 18, dtype=torch.float32, device=device)
            pad[:rows.shape[0]] = rows
            allrows = torch.empty(self.world, cap, 18, dtype=torch.float32, device=device)
            dist.all_gather_into_tensor(allrows.view(-1), pad.view(-1), group=self.group)
            rows = torch.cat([allrows[r, :counts[r]] for r in range(self.world)], dim=0)
        out = BoxList(rows[:, 0:4].contiguous(), image_size)
        out.add_field("scores", rows[:, 4].contiguous())
        out.add_field("labels", rows[:, 5].contiguous().view(torch.int32).to(torch.long))
        out.add_field("default_boxes", BoxList(rows[:, 6:10].contiguous(), image_size))
        out.add_field("transform_corners", rows[:, 10:18].contiguous())
        return out

    __call__ = forward
