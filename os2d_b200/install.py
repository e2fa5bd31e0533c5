"""Route an UNMODIFIED reference checkout through this package.

    import os2d_b200.install as hook; hook.install()      # before build_os2d_from_config() runs

`os2d/modeling/model.py:19` binds ``build_os2d_head_creator`` by name at import time and `model.py:35` constructs
``Os2dBoxCoder``; ``install()`` rebinds those names (and the head classes in ``os2d.modeling.head``) to the B200
implementations, replaces ``Os2dBoxCoder.decode_pyramid`` by an adapter around ``os2d_b200.box_coder`` and
``os2d.data.voc_eval.do_voc_evaluation`` by the on-device evaluation.  The
reference's own value types (FeatureMapSize, BoxList) are used inside this package afterwards so that objects crossing
the boundary in either direction compare equal.  ``main.py``, ``config.py`` and ``evaluate.py`` stay byte-identical.
See INTEGRATION.md for the launcher one-liner.
"""
import importlib

from . import box_coder as _bc
from . import dist as _dist
from . import evaluate as _evaluate
from . import head as _head
from . import model as _model
from . import structures as _st

_installed = False


def install(reference_package="os2d"):
    global _installed
    ref_head = importlib.import_module(reference_package + ".modeling.head")
    ref_model = importlib.import_module(reference_package + ".modeling.model")
    ref_bc = importlib.import_module(reference_package + ".modeling.box_coder")
    ref_fm = importlib.import_module(reference_package + ".structures.feature_map")
    ref_bb = importlib.import_module(reference_package + ".structures.bounding_box")

    # 1. value types: use the reference's classes inside this package
    for mod in (_st, _bc, _head, _model, _dist, _evaluate):
        if hasattr(mod, "FeatureMapSize"):
            setattr(mod, "FeatureMapSize", ref_fm.FeatureMapSize)
        if hasattr(mod, "BoxList"):
            setattr(mod, "BoxList", ref_bb.BoxList)
        if hasattr(mod, "cat_boxlist"):
            setattr(mod, "cat_boxlist", ref_bb.cat_boxlist)

    # 2. head: factory + classes (model.py:19,153; evaluate.py:273 goes through the creator instance)
    for name in ("build_os2d_head_creator", "Os2dHeadCreator", "Os2dHead", "Os2dAlignment", "TransformationNet"):
        setattr(ref_head, name, getattr(_head, name))
    ref_model.build_os2d_head_creator = _head.build_os2d_head_creator

    # 3. decode + NMS (box_coder.py:448, bounding_box.py:344)
    def decode_pyramid(self, loc_scores_pyramid, cls_scores_pyramid, img_size_pyramid, class_ids,
                       nms_score_threshold=0.0, nms_iou_threshold=0.3, inverse_box_transforms=None,
                       transform_corners_pyramid=None):
        impl = getattr(self, "_os2d_b200_impl", None)
        if impl is None:
            impl = _bc.Os2dBoxCoder(self.positive_iou_threshold, self.negative_iou_threshold,
                                    self.remap_classification_targets_iou_pos, self.remap_classification_targets_iou_neg,
                                    self.output_box_grid_generator, self.get_feature_map_size,
                                    do_nms_across_classes=self.do_nms_across_classes)
            self._os2d_b200_impl = impl
        return impl.decode_pyramid(loc_scores_pyramid, cls_scores_pyramid, img_size_pyramid, class_ids,
                                   nms_score_threshold=nms_score_threshold, nms_iou_threshold=nms_iou_threshold,
                                   inverse_box_transforms=inverse_box_transforms,
                                   transform_corners_pyramid=transform_corners_pyramid)

    ref_bc.Os2dBoxCoder.decode_pyramid = decode_pyramid
    ref_bb.nms = _bc.nms
    ref_bc.nms = _bc.nms

    # 4. detection evaluation on the device (voc_eval.py:14; evaluate.py:154 binds the name at import)
    from . import voc_eval as _voc
    ref_voc = importlib.import_module(reference_package + ".data.voc_eval")
    ref_voc.do_voc_evaluation = _voc.do_voc_evaluation

    # 5. batched-class evaluation iterator (evaluate.py:177; needs matplotlib/yacs importable, so only when it imports)
    try:
        ref_eval = importlib.import_module(reference_package + ".engine.evaluate")
        from . import evaluate as _ev
        ref_eval.make_iterator_extract_scores_from_images_batched = _ev.make_iterator_extract_scores_from_images_batched
        ref_eval.do_voc_evaluation = _voc.do_voc_evaluation
    except Exception:   # noqa: BLE001  - the reference's eval module is not importable without matplotlib / yacs
        pass
    _installed = True
    return True


def is_installed():
    return _installed
