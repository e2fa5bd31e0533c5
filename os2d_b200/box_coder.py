"""Anchor grids, box decoding and per-class NMS across pyramid levels on the GPU.

Drop-in for the inference part of os2d/modeling/box_coder.py (BoxGridGenerator :63-76,
Os2dBoxCoder.decode_pyramid :448-536, _nms_box_lists :424-437, build_loc_targets :306-317) and of
os2d/structures/bounding_box.py nms :344-387.  Training-only members (encode, remap_anchor_targets,
get_box_to_cut_anchor) are out of scope of the hot path and not provided.

The reference loops in Python over classes x levels with tiny kernels; here ``decode_pyramid`` is two launches for any
number of classes, levels and candidates (csrc/detect.cu: one CTA per real label decodes, filters, orders and runs the
chunk-of-10000 / iterate-to-fixpoint NMS of the reference exactly; a second kernel writes the survivors).  The staged
round-1 form (decode kernel per level + torch ordering + segment NMS kernel per pass) stays as ``decode_pyramid_staged``
(ablation, independent check) and behind ``nms()`` for plain BoxLists.
"""
import ctypes
import math

import torch

from . import _cabi
from .structures import BoxList, FeatureMapSize, cat_boxlist

BOX_ENCODING_WEIGHTS = torch.tensor([10, 10, 5, 5])
NMS_MAX_BATCH = 10000


class BoxGridGenerator:
    """Anchor grid specialised to a box size / stride (box_coder.py:63-76).  Boxes are row-major over the
    feature map (index y * w + x) despite the reference's 'columnfirst' name (box_coder.py:48-51)."""

    def __init__(self, box_size, box_stride):
        self.box_size = box_size
        self.box_stride = box_stride
        self._cache = {}

    def create_strided_boxes_columnfirst(self, fm_size):
        key = (fm_size.w, fm_size.h)
        if key not in self._cache:
            cx = (torch.arange(0, fm_size.w, dtype=torch.float) + 0.5) * self.box_stride.w
            cy = (torch.arange(0, fm_size.h, dtype=torch.float) + 0.5) * self.box_stride.h
            cx = cx.unsqueeze(0).expand(fm_size.h, -1).reshape(-1)
            cy = cy.unsqueeze(1).expand(-1, fm_size.w).reshape(-1)
            hw, hh = self.box_size.w / 2, self.box_size.h / 2
            self._cache[key] = torch.stack([cx - hw, cy - hh, cx + hw, cy + hh], dim=1)
        return self._cache[key]


@_cabi.on_device_of
def nms(boxes, nms_iou_threshold, nms_max_batch=NMS_MAX_BATCH, nms_score_threshold=float("-inf"),
        do_separate_per_label=False):
    """Chunked, iterated greedy NMS with the reference semantics (bounding_box.py:344-387); returns the indices
    of the surviving boxes (int64, in the reference's order: chunk order, score-descending inside a chunk)."""
    if nms_max_batch != NMS_MAX_BATCH:
        raise NotImplementedError("the NMS kernel is built for the reference chunk size of 10000 boxes")
    scores = boxes.get_field("scores")
    xyxy = boxes.bbox_xyxy
    if xyxy.device.type != "cuda":
        raise RuntimeError("os2d_b200 requires CUDA tensors (no CPU path)")
    if do_separate_per_label:
        labels = boxes.get_field("labels")
        seg_ids = labels
    else:
        seg_ids = torch.zeros(len(boxes), dtype=torch.long, device=xyxy.device)
    cand = torch.nonzero(scores > nms_score_threshold).squeeze(1)
    if do_separate_per_label:
        # the reference iterates labels.unique() (ascending) and concatenates per-label survivors
        key = seg_ids[cand]
        cand = cand[torch.sort(key, stable=True)[1]]
        uniq, counts = torch.unique_consecutive(seg_ids[cand], return_counts=True)
        counts = counts.tolist()
    else:
        counts = [cand.numel()]
    return _segmented_chunked_nms(xyxy.contiguous(), scores.float().contiguous(), cand, counts, nms_iou_threshold)


def _segmented_chunked_nms(xyxy, scores, cand, counts, iou_thr):
    """cand: candidate box indices, grouped label by label (counts per label).  Implements, for every label at
    once, the loop of bounding_box.py:356-374: split the label's survivors (in order) into chunks of 10000, run
    greedy NMS per chunk, concatenate; stop when a label has <= 1 chunk or nothing was removed."""
    lib = _cabi.load()
    dev = xyxy.device
    n_labels = len(counts)
    if n_labels > 0 and max(counts) <= NMS_MAX_BATCH:
        # every label is a single chunk: one launch, no per-label host loop.  Segment = label; order inside a segment =
        # score descending, ties by candidate position (one stable sort on the composite key (label, -score)).
        total = cand.numel()
        if total == 0:
            return cand
        counts_t = torch.tensor(counts, dtype=torch.int64)
        seg_off = torch.zeros(n_labels + 1, dtype=torch.int32)
        seg_off[1:] = counts_t.cumsum(0)
        seg_id = torch.repeat_interleave(torch.arange(n_labels), counts_t).to(dev)
        sc = scores[cand]
        # monotone map float32 -> int64 (larger score => larger value), then key = label * 2^33 - value
        bits = sc.view(torch.int32).to(torch.int64)
        mono = torch.where(bits >= 0, bits, (-2 ** 31) - bits - 1)      # negative floats: order reverses with the magnitude bits
        mono = torch.where(sc == 0, torch.zeros_like(mono), mono)        # -0.0 ties with +0.0
        key = seg_id * (2 ** 33) - mono
        perm = torch.sort(key, stable=True)[1]
        order = cand[perm].to(torch.int32).contiguous()
        keep = torch.empty(total, dtype=torch.uint8, device=dev)
        seg_off_dev = seg_off.to(dev)        # named: device temporaries must outlive the asynchronous launch
        rc = lib.os2d_nms_segments(_cabi.ptr(xyxy), _cabi.ptr(order), _cabi.ptr(seg_off_dev), n_labels,
                                   float(iou_thr), _cabi.ptr(keep), _cabi.stream_ptr())
        _cabi.check(rc, "os2d_nms_segments")
        return cand[perm][keep.bool()]
    active = [True] * n_labels
    result_per_label = [None] * n_labels
    # current survivors per label (tensor of box indices, ordered)
    cur = list(torch.split(cand, counts)) if cand.numel() > 0 else [cand.new_zeros(0) for _ in counts]
    while any(active):
        seg_lens, seg_owner = [], []
        parts = []
        for li in range(n_labels):
            if not active[li]:
                continue
            n = cur[li].numel()
            if n == 0:
                active[li] = False
                result_per_label[li] = cur[li]
                continue
            for s in range(0, n, NMS_MAX_BATCH):
                seg_lens.append(min(NMS_MAX_BATCH, n - s))
                seg_owner.append(li)
            parts.append(cur[li])
        if not parts:
            break
        ids = torch.cat(parts)                                   # concatenated candidates of the active labels
        total = ids.numel()
        seg_off = torch.tensor([0] + list(torch.tensor(seg_lens).cumsum(0).tolist()), dtype=torch.int32)
        seg_id = torch.repeat_interleave(torch.arange(len(seg_lens)), torch.tensor(seg_lens)).to(dev)
        # order inside each segment: score descending, ties by position (stable), as torchvision's CPU kernel
        sc = scores[ids]
        o1 = torch.sort(sc, descending=True, stable=True)[1]
        o2 = torch.sort(seg_id[o1], stable=True)[1]
        perm = o1[o2]                                            # positions in `ids`, grouped by segment
        order = ids[perm].to(torch.int32).contiguous()
        keep = torch.empty(total, dtype=torch.uint8, device=dev)
        seg_off_d = seg_off.to(dev)
        rc = lib.os2d_nms_segments(_cabi.ptr(xyxy), _cabi.ptr(order), _cabi.ptr(seg_off_d), len(seg_lens),
                                   float(iou_thr), _cabi.ptr(keep), _cabi.stream_ptr())
        _cabi.check(rc, "os2d_nms_segments")
        kept_sorted = ids[perm][keep.bool()]                     # survivors, segment by segment, score-descending
        kept_seg = seg_id[o1][o2][keep.bool()]
        kept_counts = torch.bincount(kept_seg, minlength=len(seg_lens)).tolist()
        # regroup per label
        pos = 0
        seg_i = 0
        for li in range(n_labels):
            if not active[li]:
                continue
            n_before = cur[li].numel()
            n_chunks = int(math.ceil(n_before / NMS_MAX_BATCH))
            n_after = sum(kept_counts[seg_i:seg_i + n_chunks])
            cur[li] = kept_sorted[pos:pos + n_after]
            pos += n_after
            seg_i += n_chunks
            if n_chunks <= 1 or n_after == n_before:
                active[li] = False
                result_per_label[li] = cur[li]
    return torch.cat(result_per_label) if n_labels > 0 else cand


class PendingDetections:
    """Handle of Os2dBoxCoder.decode_pyramid_async: ``result()`` (once) finishes the call and returns the BoxList."""

    def __init__(self, finish):
        self._finish, self._out = finish, None

    def result(self):
        if self._finish is not None:
            self._out = self._finish()
            self._finish = None
        return self._out


class Os2dBoxCoder:
    """Inference side of the reference box coder (box_coder.py:169-189, 448-536): anchors from the image-level
    box grid generator and the network's feature-map-size function, decode + NMS across the pyramid."""

    def __init__(self, positive_iou_threshold, negative_iou_threshold, remap_classification_targets_iou_pos,
                 remap_classification_targets_iou_neg, output_box_grid_generator, function_get_feature_map_size,
                 do_nms_across_classes=False):
        self.get_feature_map_size = function_get_feature_map_size
        self.output_box_grid_generator = output_box_grid_generator
        self.positive_iou_threshold = positive_iou_threshold
        self.negative_iou_threshold = negative_iou_threshold
        self.remap_classification_targets_iou_pos = remap_classification_targets_iou_pos
        self.remap_classification_targets_iou_neg = remap_classification_targets_iou_neg
        self.do_nms_across_classes = do_nms_across_classes
        self.weights = BOX_ENCODING_WEIGHTS
        self._fm_size_cache = {}

    def _get_feature_map_size_per_image_size(self, img_size):
        key = (img_size.w, img_size.h)
        if key not in self._fm_size_cache:
            self._fm_size_cache[key] = self.get_feature_map_size(img_size)
        return self._fm_size_cache[key]

    def _get_default_boxes(self, img_size):
        fm = self._get_feature_map_size_per_image_size(img_size)
        return BoxList(self.output_box_grid_generator.create_strided_boxes_columnfirst(fm), image_size=img_size, mode="xyxy")

    @staticmethod
    def build_loc_targets(class_boxes, default_boxes):
        """Box encoding used by the head (box_coder.py:306-317); fused into csrc/resample.cu on the hot path,
        kept here as a small tensor utility with the same arithmetic."""
        def clip_min(b):
            b = b.clone()
            m = (b[:, 0] + 1) > b[:, 2]
            b[m, 2] = b[m, 0] + 1
            m = (b[:, 1] + 1) > b[:, 3]
            b[m, 3] = b[m, 1] + 1
            return b
        g, a = clip_min(class_boxes.bbox_xyxy), clip_min(default_boxes.bbox_xyxy)
        aw, ah = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
        gw, gh = g[:, 2] - g[:, 0], g[:, 3] - g[:, 1]
        return torch.stack([10 * ((g[:, 0] + 0.5 * gw) - (a[:, 0] + 0.5 * aw)) / aw,
                            10 * ((g[:, 1] + 0.5 * gh) - (a[:, 1] + 0.5 * ah)) / ah,
                            5 * torch.log(gw / aw), 5 * torch.log(gh / ah)], dim=1)

    @staticmethod
    def _nms_box_lists(boxlists, nms_iou_threshold):
        """Joint NMS of several BoxLists, survivors sorted by score (box_coder.py:424-437)."""
        boxes = cat_boxlist(boxlists)
        keep = nms(boxes, nms_iou_threshold)
        sc = boxes.get_field("scores")[keep]
        keep = keep[torch.sort(sc, dim=0, descending=True)[1]]
        return boxes[keep]

    def decode_pyramid(self, loc_scores_pyramid, cls_scores_pyramid, img_size_pyramid, class_ids,
                       nms_score_threshold=0.0, nms_iou_threshold=0.3, inverse_box_transforms=None,
                       transform_corners_pyramid=None):
        """Same contract as box_coder.py:448-536 (see decode_pyramid_async, of which this is the synchronous form)."""
        return self.decode_pyramid_async(loc_scores_pyramid, cls_scores_pyramid, img_size_pyramid, class_ids,
                                         nms_score_threshold=nms_score_threshold, nms_iou_threshold=nms_iou_threshold,
                                         inverse_box_transforms=inverse_box_transforms,
                                         transform_corners_pyramid=transform_corners_pyramid).result()

    @_cabi.on_device_of
    def decode_pyramid_async(self, loc_scores_pyramid, cls_scores_pyramid, img_size_pyramid, class_ids,
                             nms_score_threshold=0.0, nms_iou_threshold=0.3, inverse_box_transforms=None,
                             transform_corners_pyramid=None):
        """Launches the decode + NMS kernel and returns a handle; ``handle.result()`` reads the detection count back (the only
        host synchronisation - it waits for THIS kernel only, work enqueued in between keeps the GPU busy), launches the
        gather kernel and returns the BoxList.  Pipelines of images call result() of image i after submitting image i+1.

        Same contract as box_coder.py:448-536.  loc [C,4,N_l], cls [C,N_l] (and corners [C,8,N_l]) per level ->
        BoxList with fields scores, labels, default_boxes (, transform_corners).  ``inverse_box_transforms`` may be
        reference TransformList objects that only resize (their effect is obtained by probing them with a
        box list) or anything with a ``target_size`` / callable returning a resized BoxList.

        Two launches for any number of classes, levels and candidates (csrc/detect.cu): os2d_detect_pyramid (one CTA per
        real label: decode, filter, candidate order, chunked NMS to the fixpoint, final order) and
        os2d_gather_detections; the only host synchronisation is the 4-byte read of the detection count between them."""
        lib = _cabi.load()
        num_classes = len(class_ids)
        device = cls_scores_pyramid[0].device
        if device.type != "cuda":
            raise RuntimeError("os2d_b200 requires CUDA tensors (no CPU path)")
        gen = self.output_box_grid_generator
        have_corners = transform_corners_pyramid is not None
        st = _cabi.stream_ptr()
        L = len(cls_scores_pyramid)
        if L > _cabi.MAX_PYRAMID_LEVELS:
            raise NotImplementedError("os2d_detect_pyramid handles up to {} pyramid levels".format(_cabi.MAX_PYRAMID_LEVELS))
        levels = (_cabi.PyramidLevel * L)()
        alive = []          # contiguous fp32 copies must outlive the asynchronous launches
        out_size = None
        sum_n = 0
        for i_p, (loc, cls) in enumerate(zip(loc_scores_pyramid, cls_scores_pyramid)):
            assert cls.device == device and loc.device == device, "scores and boxes should be on the same device"
            img_size = img_size_pyramid[i_p]
            fm = self._get_feature_map_size_per_image_size(img_size)
            N = fm.w * fm.h
            assert cls.shape == (num_classes, N) and loc.shape == (num_classes, 4, N)
            if inverse_box_transforms is not None:
                target = _probe_transform_target(inverse_box_transforms[i_p], img_size)
                rw, rh = float(target.w) / img_size.w, float(target.h) / img_size.h
            else:
                target, rw, rh = img_size, 1.0, 1.0
            if out_size is None:
                out_size = target
            assert target == out_size, "all pyramid levels must map to the same image size (bounding_box.py:403)"
            loc_c = loc.float().contiguous()
            cls_c = cls.float().contiguous()
            cor_c = transform_corners_pyramid[i_p].float().contiguous() if have_corners else None
            if have_corners:
                assert cor_c.shape == (num_classes, 8, N)
            alive += [loc_c, cls_c, cor_c]
            lv = levels[i_p]
            lv.loc, lv.score = loc_c.data_ptr(), cls_c.data_ptr()
            lv.corners = cor_c.data_ptr() if have_corners else None
            lv.num_anchors, lv.fm_w = N, fm.w
            lv.img_w, lv.img_h = float(img_size.w), float(img_size.h)
            lv.scale_x, lv.scale_y, lv.same_scale = rw, rh, 1 if rw == rh else 0
            sum_n += N

        view_off, view_ids, label_values, n_labels, max_views = self._label_tables(class_ids, device)
        F = num_classes * sum_n
        cand = torch.empty(F, dtype=torch.int32, device=device)
        out_ids = torch.empty(F, dtype=torch.int32, device=device)
        keys = torch.empty(F, dtype=torch.int64, device=device) if max_views * sum_n > NMS_MAX_BATCH else None
        counts = torch.empty(n_labels, dtype=torch.int32, device=device)
        offsets = torch.empty(n_labels + 1, dtype=torch.int32, device=device)
        done = self._done_counter(device)
        grid_args = (float(gen.box_stride.w), float(gen.box_stride.h), float(gen.box_size.w), float(gen.box_size.h))
        rc = lib.os2d_detect_pyramid(levels, L, num_classes, _cabi.ptr(view_off), _cabi.ptr(view_ids), n_labels, max_views,
                                     *grid_args, float(nms_score_threshold), float(nms_iou_threshold), _cabi.ptr(cand),
                                     _cabi.ptr(keys), _cabi.ptr(out_ids), _cabi.ptr(counts), _cabi.ptr(offsets),
                                     _cabi.ptr(done), st)
        _cabi.check(rc, "os2d_detect_pyramid")
        count_host = self._pinned_count()
        count_host.copy_(offsets[n_labels:n_labels + 1], non_blocking=True)
        counted = torch.cuda.Event()
        stream = torch.cuda.current_stream()
        counted.record(stream)
        out_size = out_size if out_size is not None else img_size_pyramid[0]

        def finish():
            counted.synchronize()                                  # the one host synchronisation: number of detections
            total = int(count_host[0])
            self.__dict__.setdefault("_count_free", []).append(count_host)       # back to the pool of pinned scalars
            with torch.cuda.device(device), torch.cuda.stream(stream):
                boxes = torch.empty(total, 4, dtype=torch.float32, device=device)
                scores = torch.empty(total, dtype=torch.float32, device=device)
                labels = torch.empty(total, dtype=torch.long, device=device)
                anchors = torch.empty(total, 4, dtype=torch.float32, device=device)
                corners = torch.empty(total, 8, dtype=torch.float32, device=device) if have_corners else None
                if total > 0:
                    rc = lib.os2d_gather_detections(levels, L, num_classes, _cabi.ptr(view_off), n_labels, *grid_args,
                                                    _cabi.ptr(out_ids), _cabi.ptr(counts), _cabi.ptr(offsets),
                                                    _cabi.ptr(label_values), _cabi.ptr(boxes), _cabi.ptr(scores), _cabi.ptr(labels),
                                                    _cabi.ptr(anchors), _cabi.ptr(corners), ctypes.c_void_p(stream.cuda_stream))
                    _cabi.check(rc, "os2d_gather_detections")
            # the inputs' fp32 copies and the workspaces (`alive`, closure) stay referenced until both launches are enqueued;
            # the caching allocator hands their blocks out again in stream order on this same stream
            alive.clear()
            out = BoxList(boxes, out_size)
            out.add_field("scores", scores)
            out.add_field("default_boxes", BoxList(anchors, out.image_size))
            out.add_field("labels", labels)
            if have_corners:
                out.add_field("transform_corners", corners)
            if self.do_nms_across_classes and len(out) > 0:
                out = self._nms_box_lists([out], nms_iou_threshold)
            return out

        return PendingDetections(finish)

    def _pinned_count(self):
        """A pinned int32 scalar for the asynchronous read-back of the detection count: taken from a pool, returned by
        ``finish`` (any number of handles may be in flight; a handle that is dropped unread just keeps its scalar)."""
        free = self.__dict__.setdefault("_count_free", [])
        return free.pop() if free else torch.empty(1, dtype=torch.int32).pin_memory()

    def _label_tables(self, class_ids, device):
        """Device tables of the label structure, cached per class-id tuple: views of real label i (set order, box_coder.py:483)
        are view_ids[view_off[i]:view_off[i+1]] in class-view order."""
        key = (tuple(int(c) for c in class_ids), device.index)
        cache = self.__dict__.setdefault("_label_cache", {})
        if key not in cache:
            label_order = list(set(class_ids))                       # same iteration order as box_coder.py:483
            views = [[i for i, c in enumerate(class_ids) if c == l] for l in label_order]
            off = [0]
            for v in views:
                off.append(off[-1] + len(v))
            if len(cache) > 64:
                cache.clear()
            cache[key] = (torch.tensor(off, dtype=torch.int32).to(device),
                          torch.tensor([i for v in views for i in v], dtype=torch.int32).to(device),
                          torch.tensor([int(l) for l in label_order], dtype=torch.long).to(device),
                          len(label_order), max(len(v) for v in views))
        return cache[key]

    def _done_counter(self, device):
        """Zero-initialised completion counter of os2d_detect_pyramid (reset by the kernel), one per device and stream."""
        key = (device.index, torch.cuda.current_stream().cuda_stream)
        cache = self.__dict__.setdefault("_done_cache", {})
        if key not in cache:
            cache[key] = torch.zeros(1, dtype=torch.int32, device=device)
        return cache[key]

    @_cabi.on_device_of
    def decode_pyramid_staged(self, loc_scores_pyramid, cls_scores_pyramid, img_size_pyramid, class_ids,
                              nms_score_threshold=0.0, nms_iou_threshold=0.3, inverse_box_transforms=None,
                              transform_corners_pyramid=None):
        """The staged form of decode_pyramid (round-1 path, kept as the unfused ablation and as an independent check of the
        fused kernels at sizes the CPU oracle cannot reach): os2d_decode_boxes per level, candidate ordering / sorting with
        torch index ops, os2d_nms_segments per pass.  Same results bit for bit (tests/test_gpu_postproc.py)."""
        lib = _cabi.load()
        num_classes = len(class_ids)
        device = cls_scores_pyramid[0].device
        if device.type != "cuda":
            raise RuntimeError("os2d_b200 requires CUDA tensors (no CPU path)")
        gen = self.output_box_grid_generator
        have_corners = transform_corners_pyramid is not None
        st = _cabi.stream_ptr()

        lvl = []   # per level: boxes [C,N,4], anchors [N,4], corners [C,N,8], valid [C,N], scores [C,N]
        out_size = None
        for i_p, (loc, cls) in enumerate(zip(loc_scores_pyramid, cls_scores_pyramid)):
            assert cls.device == device and loc.device == device, "scores and boxes should be on the same device"
            img_size = img_size_pyramid[i_p]
            fm = self._get_feature_map_size_per_image_size(img_size)
            N = fm.w * fm.h
            assert cls.shape == (num_classes, N) and loc.shape == (num_classes, 4, N)
            if inverse_box_transforms is not None:
                target = _probe_transform_target(inverse_box_transforms[i_p], img_size)
                rw, rh = float(target.w) / img_size.w, float(target.h) / img_size.h
            else:
                target, rw, rh = img_size, 1.0, 1.0
            if out_size is None:
                out_size = target
            assert target == out_size, "all pyramid levels must map to the same image size (bounding_box.py:403)"
            loc_c = loc.float().contiguous()
            cls_c = cls.float().contiguous()
            cor_c = transform_corners_pyramid[i_p].float().contiguous() if have_corners else None
            boxes = torch.empty(num_classes, N, 4, dtype=torch.float32, device=device)
            anchors = torch.empty(N, 4, dtype=torch.float32, device=device)
            cor_out = torch.empty(num_classes, N, 8, dtype=torch.float32, device=device) if have_corners else None
            valid = torch.empty(num_classes, N, dtype=torch.uint8, device=device)
            rc = lib.os2d_decode_boxes(num_classes, N, fm.w, float(gen.box_stride.w), float(gen.box_stride.h),
                                       float(gen.box_size.w), float(gen.box_size.h), float(img_size.w), float(img_size.h),
                                       float(nms_score_threshold), rw, rh, 1 if rw == rh else 0, _cabi.ptr(loc_c),
                                       _cabi.ptr(cls_c), _cabi.ptr(cor_c), _cabi.ptr(boxes), _cabi.ptr(anchors),
                                       _cabi.ptr(cor_out), _cabi.ptr(valid), st)
            _cabi.check(rc, "os2d_decode_boxes")
            lvl.append((boxes, anchors, cor_out, valid, cls_c, N))

        # ---- candidate list in the reference's concatenation order: label (set order), class view, level, anchor ----
        # Candidates are FLAT indices into the per-level arrays concatenated over levels: f = base_l + c * N_l + n, so that
        # the NMS kernel reads boxes in place and only the survivors are gathered.
        label_order = list(set(class_ids))                       # same iteration order as box_coder.py:483
        label_rank = {l: r for r, l in enumerate(label_order)}
        n_labels = len(label_order)
        ranks_host = [label_rank[c] for c in class_ids]
        L = len(lvl)
        rank_of_view = torch.tensor(ranks_host, device=device)
        n_per_level = [t[5] for t in lvl]
        bases = [0]
        for nl in n_per_level:
            bases.append(bases[-1] + num_classes * nl)
        cat = (lambda xs: xs[0]) if L == 1 else (lambda xs: torch.cat(xs, dim=0))
        boxes_flat = cat([t[0].view(-1, 4) for t in lvl])
        scores_flat = cat([t[4].view(-1) for t in lvl])
        valid_flat = cat([t[3].view(-1) for t in lvl])
        cand = torch.nonzero(valid_flat).squeeze(1)              # ascending: level, class view, anchor

        def split_index(f):
            """flat index -> (level, class view, anchor)"""
            if L == 1:
                return torch.zeros_like(f), f // n_per_level[0], f % n_per_level[0]
            lv = torch.bucketize(f, torch.tensor(bases[1:-1], device=device), right=True)
            base_t = torch.tensor(bases[:-1], device=device)[lv]
            nl_t = torch.tensor(n_per_level, device=device)[lv]
            return lv, (f - base_t) // nl_t, (f - base_t) % nl_t

        c_lvl, c_cls, c_n = split_index(cand)
        c_rank = rank_of_view[c_cls]
        if L > 1 or any(ranks_host[i] > ranks_host[i + 1] for i in range(len(ranks_host) - 1)):
            key = (c_rank * num_classes + c_cls) * L + c_lvl
            perm = torch.sort(key, stable=True)[1]
            cand, c_rank = cand[perm], c_rank[perm]

        # static bound: candidates of a label <= (#views of the label) * (anchors of all levels)
        views_per_label = max(ranks_host.count(r) for r in set(ranks_host))
        if views_per_label * sum(n_per_level) <= NMS_MAX_BATCH:
            # every label is one NMS chunk: stays on the device (no host synchronisation), survivors leave the kernel
            # grouped by label in set order and score-descending, which is the reference's output order
            keep = _single_chunk_nms_device(boxes_flat, scores_flat, cand, c_rank, n_labels, nms_iou_threshold)
        else:
            counts = [c for c in torch.bincount(c_rank, minlength=n_labels).tolist() if c > 0]
            keep = _segmented_chunked_nms(boxes_flat, scores_flat, cand, counts, nms_iou_threshold)
            if counts and max(counts) > NMS_MAX_BATCH:
                # chunked path: per label, sort the survivors by score, descending (box_coder.py:431-435)
                k_rank = rank_of_view[split_index(keep)[1]]
                o1 = torch.sort(scores_flat[keep], descending=True, stable=True)[1]
                o2 = torch.sort(k_rank[o1], stable=True)[1]
                keep = keep[o1[o2]]

        k_lvl, k_cls, k_n = split_index(keep)
        out = BoxList(boxes_flat[keep], out_size if out_size is not None else img_size_pyramid[0])
        out.add_field("scores", scores_flat[keep])
        labels_t = torch.tensor(label_order, dtype=torch.long, device=device)
        anchors_all = cat([t[1] for t in lvl])                   # [sum N_l, 4]
        anchor_base = torch.tensor([sum(n_per_level[:i]) for i in range(L)], device=device)
        out.add_field("default_boxes", BoxList(anchors_all[anchor_base[k_lvl] + k_n], out.image_size))
        out.add_field("labels", labels_t[rank_of_view[k_cls]] if keep.numel() > 0 else torch.zeros(0, dtype=torch.long, device=device))
        if have_corners:
            out.add_field("transform_corners", cat([t[2].view(-1, 8) for t in lvl])[keep])
        if self.do_nms_across_classes and len(out) > 0:
            out = self._nms_box_lists([out], nms_iou_threshold)
        return out


def _single_chunk_nms_device(xyxy, scores, cand, seg_of_cand, n_segs, iou_thr):
    """Greedy NMS of every segment (each known to hold <= 10000 candidates) in ONE launch without touching the host:
    segment offsets come from a device-side bincount/cumsum; order inside a segment = score descending, ties by candidate
    position (one stable sort on the composite key (segment, -score)).  Returns the surviving candidate ids grouped by
    segment, score-descending."""
    lib = _cabi.load()
    dev = xyxy.device
    total = cand.numel()
    if total == 0:
        return cand
    seg_off = torch.zeros(n_segs + 1, dtype=torch.int32, device=dev)
    seg_off[1:] = torch.bincount(seg_of_cand, minlength=n_segs).cumsum(0)
    sc = scores[cand]
    bits = sc.view(torch.int32).to(torch.int64)
    mono = torch.where(bits >= 0, bits, (-2 ** 31) - bits - 1)      # monotone float32 -> int64 map
    mono = torch.where(sc == 0, torch.zeros_like(mono), mono)        # -0.0 ties with +0.0
    perm = torch.sort(seg_of_cand * (2 ** 33) - mono, stable=True)[1]
    ordered = cand[perm]
    order32 = ordered.to(torch.int32)
    keep = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    rc = lib.os2d_nms_segments(_cabi.ptr(xyxy), _cabi.ptr(order32), _cabi.ptr(seg_off), n_segs, float(iou_thr),
                               _cabi.ptr(keep), _cabi.stream_ptr())
    _cabi.check(rc, "os2d_nms_segments")
    return ordered[keep[:total].bool()]


class _Resize:
    """Minimal inverse box transform: rescale to ``target_size`` (what the reference TransformList built by the
    eval dataloader does to boxes, os2d/structures/transforms.py:12-52 -> BoxList.resize)."""

    def __init__(self, target_size):
        self.target_size = target_size

    def __call__(self, boxes):
        return boxes.resize(self.target_size)


def make_resize_transform(target_size):
    return _Resize(target_size)


def _probe_transform_target(transform, img_size):
    """Image size a box transform maps ``img_size`` boxes to."""
    if hasattr(transform, "target_size"):
        return transform.target_size
    # probe with a non-degenerate box: the decode kernel applies the inverse transform as a per-axis rescale
    # (BoxList.resize, bounding_box.py:138-163), so anything else (flips, crops) must be refused, not mis-placed
    box = torch.tensor([[1.0, 2.0, 3.0, 5.0]])
    res = transform(BoxList(box.clone(), img_size))
    target = FeatureMapSize(w=res.image_size.w, h=res.image_size.h)
    rw, rh = float(target.w) / img_size.w, float(target.h) / img_size.h
    expect = box * (torch.tensor([rw, rw, rw, rw]) if rw == rh else torch.tensor([rw, rh, rw, rh]))
    if not torch.allclose(res.bbox_xyxy.float().cpu(), expect, rtol=1e-5, atol=1e-5):
        raise NotImplementedError("os2d_b200.decode_pyramid supports inverse box transforms that only resize "
                                  "(got a transform that maps {} to {})".format(box.tolist(), res.bbox_xyxy.tolist()))
    return target
