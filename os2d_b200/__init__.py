"""os2d_b200 - B200-native (sm_100a) implementation of the OS2D dense correlation-and-alignment head.

Public surface mirrors the reference modules for the hot path:
    os2d_b200.head       <-> os2d/modeling/head.py
    os2d_b200.box_coder  <-> os2d/modeling/box_coder.py (inference part) + bounding_box.nms
    os2d_b200.structures <-> os2d/structures/{feature_map,bounding_box}.py
    os2d_b200.model      <-> os2d/modeling/model.py (forward / apply_class_heads)
    os2d_b200.evaluate   <-> os2d/engine/evaluate.py make_iterator_extract_scores_from_images_batched (batched classes)
    os2d_b200.dist       class-axis sharding over GPUs + all-gather of per-class outputs
    os2d_b200.install    monkey-patch hook that routes an unmodified reference main.py through these classes
"""
from .structures import FeatureMapSize, BoxList, cat_boxlist  # noqa: F401
from .box_coder import BoxGridGenerator, Os2dBoxCoder, nms, make_resize_transform  # noqa: F401
from .head import (build_os2d_head_creator, Os2dAlignment, Os2dHeadCreator, Os2dHead, TransformationNet,  # noqa: F401
                   normalize_feature_map_L2)
from .graphed import GraphedHead  # noqa: F401

__all__ = ["FeatureMapSize", "BoxList", "cat_boxlist", "BoxGridGenerator", "Os2dBoxCoder", "nms",
           "make_resize_transform", "build_os2d_head_creator", "Os2dAlignment", "Os2dHeadCreator", "Os2dHead",
           "TransformationNet", "normalize_feature_map_L2", "GraphedHead"]
