"""Anchor grids, box decoding and per-class NMS across pyramid levels on the GPU.

Drop-in for the inference part of os2d/modeling/box_coder.py (BoxGridGenerator :63-76,
Os2dBoxCoder.decode_pyramid :448-536, _nms_box_lists :424-437, build_loc_targets :306-317) and of
os2d/structures/bounding_box.py nms :344-387.  Training-only members (encode, remap_anchor_targets,
get_box_to_cut_anchor) are out of scope of the hot path and not provided.

The reference loops in Python over classes x levels with tiny kernels; here one decode launch handles all
classes of a level (csrc/postproc.cu decode_kernel) and one NMS launch handles all (label, chunk) segments
(nms_kernel), reproducing the chunk-of-10000 / iterate-to-fixpoint semantics of the reference exactly.
"""
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
        """Same contract as box_coder.py:448-536.  loc [C,4,N_l], cls [C,N_l] (and corners [C,8,N_l]) per level ->
        BoxList with fields scores, labels, default_boxes (, transform_corners).  ``inverse_box_transforms`` may be
        reference TransformList objects that only resize (their effect is obtained by probing them with a unit
        box list) or anything with a ``target_size`` / callable returning a resized BoxList."""
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
        label_order = list(set(class_ids))                       # same iteration order as box_coder.py:483
        label_rank = {l: r for r, l in enumerate(label_order)}
        L = len(lvl)
        rank_of_view = torch.tensor([label_rank[c] for c in class_ids], device=device)
        cand_cls, cand_lvl, cand_n = [], [], []
        for i_p, (boxes, anchors, cor_out, valid, cls_c, N) in enumerate(lvl):
            nz = torch.nonzero(valid)                            # sorted by class view, then anchor
            cand_cls.append(nz[:, 0])
            cand_n.append(nz[:, 1])
            cand_lvl.append(torch.full_like(nz[:, 0], i_p))
        cand_cls = torch.cat(cand_cls)
        cand_n = torch.cat(cand_n)
        cand_lvl = torch.cat(cand_lvl)
        key = (rank_of_view[cand_cls] * num_classes + cand_cls) * L + cand_lvl
        perm = torch.sort(key, stable=True)[1]
        cand_cls, cand_n, cand_lvl = cand_cls[perm], cand_n[perm], cand_lvl[perm]
        M = cand_cls.numel()

        # gather candidate boxes / fields
        all_boxes = torch.empty(M, 4, dtype=torch.float32, device=device)
        all_scores = torch.empty(M, dtype=torch.float32, device=device)
        all_anchors = torch.empty(M, 4, dtype=torch.float32, device=device)
        all_corners = torch.empty(M, 8, dtype=torch.float32, device=device) if have_corners else None
        for i_p, (boxes, anchors, cor_out, valid, cls_c, N) in enumerate(lvl):
            m = cand_lvl == i_p
            if L == 1:
                m = slice(None)
            c_i, n_i = cand_cls[m], cand_n[m]
            all_boxes[m] = boxes[c_i, n_i]
            all_scores[m] = cls_c[c_i, n_i]
            all_anchors[m] = anchors[n_i]
            if have_corners:
                all_corners[m] = cor_out[c_i, n_i]
        labels_t = torch.tensor(label_order, dtype=torch.long, device=device)
        cand_rank = rank_of_view[cand_cls]
        all_labels = labels_t[cand_rank] if M > 0 else torch.zeros(0, dtype=torch.long, device=device)

        if M == 0:
            # the reference would fail in cat_boxlist([]) (box_coder.py:534); return an empty list instead
            out = BoxList(all_boxes, out_size if out_size is not None else img_size_pyramid[0])
        else:
            counts = torch.bincount(cand_rank, minlength=len(label_order)).tolist()
            counts = [c for c in counts if c > 0]
            cand = torch.arange(M, device=device)
            keep = _segmented_chunked_nms(all_boxes, all_scores, cand, counts, nms_iou_threshold)
            # per label: sort survivors by score, descending (box_coder.py:431-435); labels stay in set order
            k_rank = cand_rank[keep]
            k_sc = all_scores[keep]
            o1 = torch.sort(k_sc, descending=True, stable=True)[1]
            o2 = torch.sort(k_rank[o1], stable=True)[1]
            keep = keep[o1[o2]]
            out = BoxList(all_boxes[keep], out_size)
            all_scores, all_labels, all_anchors = all_scores[keep], all_labels[keep], all_anchors[keep]
            if have_corners:
                all_corners = all_corners[keep]
        out.add_field("scores", all_scores)
        out.add_field("default_boxes", BoxList(all_anchors, out.image_size))
        out.add_field("labels", all_labels)
        if have_corners:
            out.add_field("transform_corners", all_corners)
        if self.do_nms_across_classes and len(out) > 0:
            out = self._nms_box_lists([out], nms_iou_threshold)
        return out


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
    probe = BoxList(torch.zeros(1, 4), img_size)
    res = transform(probe)
    return FeatureMapSize(w=res.image_size.w, h=res.image_size.h)
