"""Value types on the boundary of the hot path: FeatureMapSize, BoxList, cat_boxlist.

Same public surface as the reference types (os2d/structures/feature_map.py:5-44,
os2d/structures/bounding_box.py:15-304, 390-418) for the members the head / decode path touches;
written from scratch.
"""
import torch


class FeatureMapSize(object):
    """Immutable (w, h) pair; from a PIL image it is (width, height), from a tensor (size(-1), size(-2)).
    Hashable (used as a cache key).  Reference: os2d/structures/feature_map.py:5-44."""
    __slots__ = ("w", "h")

    def __init__(self, img=None, w=None, h=None):
        if w is None or h is None:
            if isinstance(img, torch.Tensor):
                w, h = img.size(-1), img.size(-2)
            elif hasattr(img, "size") and not callable(img.size):   # PIL.Image.Image
                w, h = img.size
            else:
                raise RuntimeError("Cannot initialize FeatureMapSize")
        object.__setattr__(self, "w", w)
        object.__setattr__(self, "h", h)

    def __setattr__(self, *args):
        raise AttributeError("Attributes of FeatureMapSize cannot be changed")

    def __delattr__(self, *args):
        raise AttributeError("Attributes of FeatureMapSize cannot be deleted")

    def __repr__(self):
        return "FeatureMapSize(w={}, h={})".format(self.w, self.h)

    def __eq__(self, other):
        return hasattr(other, "w") and hasattr(other, "h") and (self.w, self.h) == (other.w, other.h)

    def __hash__(self):
        return hash((self.w, self.h))


class BoxList(object):
    """xyxy boxes [n,4] + per-box fields, tagged with an image size.
    Reference: os2d/structures/bounding_box.py:15-304 (subset used by head / decode / nms)."""

    def __init__(self, bbox, image_size, mode="xyxy"):
        if not isinstance(bbox, torch.Tensor):
            raise ValueError("bbox should be of type torch.Tensor")
        bbox = bbox.to(dtype=torch.float32)
        if bbox.ndimension() != 2 or bbox.size(-1) != 4:
            raise ValueError("bbox should be of size n x 4, got {}".format(tuple(bbox.shape)))
        if mode == "xyxy":
            pass
        elif mode == "xywh":
            bbox = torch.cat([bbox[:, :2], bbox[:, :2] + bbox[:, 2:]], dim=-1)
        elif mode == "cx_cy_w_h":
            bbox = torch.cat([bbox[:, :2] - bbox[:, 2:] / 2, bbox[:, :2] + bbox[:, 2:] / 2], dim=-1)
        else:
            raise ValueError("mode should be xyxy, xywh or cx_cy_w_h")
        self.bbox_xyxy = bbox
        self.image_size = image_size
        self.extra_fields = {}

    @staticmethod
    def create_empty(image_size):
        return BoxList(torch.zeros(0, 4), image_size)

    def __len__(self):
        return self.bbox_xyxy.shape[0]

    def add_field(self, field, field_data):
        self.extra_fields[field] = field_data

    def get_field(self, field):
        return self.extra_fields[field]

    def has_field(self, field):
        return field in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def remove_field(self, field):
        if field not in self.extra_fields:
            raise ValueError("bbox has not field {}".format(field))
        del self.extra_fields[field]

    def _map_fields(self, fn):
        out = BoxList(fn(self.bbox_xyxy), self.image_size)
        for k, v in self.extra_fields.items():
            out.add_field(k, v._map_fields(fn) if isinstance(v, BoxList) else (fn(v) if isinstance(v, torch.Tensor) else v))
        return out

    def to(self, device):
        return self._map_fields(lambda t: t.to(device))

    def cuda(self):
        return self._map_fields(lambda t: t.cuda())

    def cpu(self):
        return self._map_fields(lambda t: t.cpu())

    def __getitem__(self, item):
        boxes = self.bbox_xyxy[item]
        if boxes.ndimension() == 1:
            boxes = boxes.view(1, 4)
        out = BoxList(boxes, self.image_size)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item])
        return out

    def resize(self, target_size):
        """Scaled copy (bounding_box.py:138-163): one multiply when both ratios agree, per axis otherwise."""
        ratio_w = float(target_size.w) / self.image_size.w
        ratio_h = float(target_size.h) / self.image_size.h
        if ratio_w == ratio_h:
            scaled = self.bbox_xyxy * ratio_w
        else:
            scale = torch.tensor([ratio_w, ratio_h, ratio_w, ratio_h], dtype=torch.float32, device=self.bbox_xyxy.device)
            scaled = self.bbox_xyxy * scale
        out = BoxList(scaled, target_size)
        for k, v in self.extra_fields.items():
            out.add_field(k, v)
        return out

    def clip_to_image(self, remove_empty=True):
        b = self.bbox_xyxy
        self.bbox_xyxy = torch.stack([b[:, 0].clamp(0, self.image_size.w), b[:, 1].clamp(0, self.image_size.h),
                                      b[:, 2].clamp(0, self.image_size.w), b[:, 3].clamp(0, self.image_size.h)], dim=1)
        if remove_empty:
            return self[~self.get_mask_empty_boxes()]
        return self

    def get_mask_empty_boxes(self):
        b = self.bbox_xyxy
        return (b[:, 3] <= b[:, 1]) | (b[:, 2] <= b[:, 0])

    def area(self):
        b = self.bbox_xyxy
        return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])

    def __repr__(self):
        return "BoxList(num_boxes={}, image_width={}, image_height={})".format(len(self), self.image_size.w,
                                                                                self.image_size.h)


def cat_boxlist(bboxes):
    """Concatenate BoxLists of the same image size and field set (bounding_box.py:390-418)."""
    assert isinstance(bboxes, (list, tuple)) and len(bboxes) > 0
    image_size = bboxes[0].image_size
    assert all(b.image_size == image_size for b in bboxes)
    fields = set(bboxes[0].fields())
    assert all(set(b.fields()) == fields for b in bboxes)
    out = BoxList(torch.cat([b.bbox_xyxy for b in bboxes], dim=0), image_size)
    for f in bboxes[0].fields():
        vals = [b.get_field(f) for b in bboxes]
        out.add_field(f, cat_boxlist(vals) if isinstance(vals[0], BoxList) else torch.cat(vals, dim=0))
    return out
