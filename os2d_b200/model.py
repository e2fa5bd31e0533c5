"""Os2dModel boundary: the reference forward API (os2d/modeling/model.py:123-288) on top of the B200 head.

The backbone (ResNet-50/101 truncated at C4, stride 16, 1024 channels) is *not* part of the hot path and
runs as the stock torchvision/cuDNN module; parameter names match the reference extractor
(``net_feature_maps.{conv1,bn1,layer1,layer2,layer3}.*``) so reference checkpoints load with
``load_state_dict``.  Only the inference regime is provided (train_mode=False).
"""
import logging
import math

import torch
import torch.nn as nn
from torchvision.models.resnet import resnet50, resnet101

from . import _cabi
from .head import build_os2d_head_creator, PackedFeatureMaps
from .structures import FeatureMapSize


class ResNetC4(nn.Module):
    """conv1..layer3 of a torchvision ResNet (reference: os2d/modeling/feature_extractor.py:23-72)."""

    def __init__(self, arch):
        super(ResNetC4, self).__init__()
        if arch.lower() == "resnet50":
            full = resnet50()
        elif arch.lower() == "resnet101":
            full = resnet101()
        else:
            raise RuntimeError("Unknown backbone arch: {0}".format(arch))
        self.conv1, self.bn1, self.relu, self.maxpool = full.conv1, full.bn1, full.relu, full.maxpool
        self.layer1, self.layer2, self.layer3 = full.layer1, full.layer2, full.layer3
        self.feature_map_stride = FeatureMapSize(h=16, w=16)
        self.feature_map_receptive_field = FeatureMapSize(h=16, w=16)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        return self.layer3(self.layer2(self.layer1(x)))

    @torch.no_grad()
    def forward_packed(self, x, half=False):
        """Channels-last run of the same network whose LAST step - the residual add + ReLU that ends layer3's final
        bottleneck (feature_extractor.py:23-72) - is fused with the head's image-side L2 normalisation (head.py:339) into
        one kernel that writes the fp16 [B, H*W, D] GEMM operand (os2d_pack_image_features_nhwc).  The fp32 NCHW feature map,
        its read-back by the head's pack kernel and the transpose never exist.  Eval mode only.  ``half=True`` additionally
        runs the convolutions in fp16 (cuDNN tensor cores): that changes the backbone's numerics, which are outside the
        parity-checked head path - the default keeps fp32."""
        if self.training:
            raise RuntimeError("forward_packed implements the evaluation regime (eval-mode BatchNorm) only")
        if x.device.type != "cuda":
            raise RuntimeError("os2d_b200 requires CUDA tensors (no CPU path)")
        with torch.cuda.device(x.device):
            ctx = torch.autocast("cuda", dtype=torch.float16) if half else torch.autocast("cuda", enabled=False)
            with ctx:
                x = x.contiguous(memory_format=torch.channels_last)
                x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
                x = self.layer2(self.layer1(x))
                for blk in list(self.layer3.children())[:-1]:
                    x = blk(x)
                blk = list(self.layer3.children())[-1]
                assert blk.downsample is None, "the last bottleneck of layer3 has an identity shortcut in ResNet-50/101"
                out = blk.relu(blk.bn1(blk.conv1(x)))
                out = blk.relu(blk.bn2(blk.conv2(out)))
                out = blk.bn3(blk.conv3(out))
            B, D, H, W = out.shape
            a = out.permute(0, 2, 3, 1)          # NHWC memory of a channels_last tensor: a contiguous [B,H,W,D] view
            b = x.permute(0, 2, 3, 1)
            if not a.is_contiguous() or not b.is_contiguous() or a.dtype != b.dtype:
                a, b = a.contiguous(), b.to(a.dtype).contiguous()
            packed = torch.empty(B, H * W, D, dtype=torch.float16, device=out.device)
            lib = _cabi.load()
            _cabi.check(lib.os2d_pack_image_features_nhwc(_cabi.ptr(a), _cabi.ptr(b), 1 if a.dtype == torch.float16 else 0, 1,
                                                          B * H * W, D, _cabi.ptr(packed), _cabi.stream_ptr()),
                        "os2d_pack_image_features_nhwc")
            cur = torch.cuda.current_stream()
            a.record_stream(cur)
            b.record_stream(cur)
        return PackedFeatureMaps(packed, H, W)


class LabelFeatureExtractor(nn.Module):
    """Runs the class-image branch on a list of differently sized images (model.py:71-95)."""

    def __init__(self, feature_extractor):
        super(LabelFeatureExtractor, self).__init__()
        self.net_class_features = feature_extractor

    def forward(self, class_image_list):
        """list of [3,h_i,w_i] images -> list of [1,D,h_i/16,w_i/16] feature maps, in order.  The reference pushes the
        images through the backbone one by one (model.py:80-88); here images of equal size form one batch (eval-mode
        BatchNorm: every sample is independent), and the per-class maps are views of the batch outputs that the ragged
        pack kernel reads in place."""
        groups = {}
        for i, img in enumerate(class_image_list):
            groups.setdefault((img.size(-2), img.size(-1)), []).append(i)
        out = [None] * len(class_image_list)
        for idx in groups.values():
            feats = self.net_class_features(torch.stack([class_image_list[i] for i in idx], dim=0))
            for j, i in enumerate(idx):
                out[i] = feats[j:j + 1]
        return out


class Os2dModel(nn.Module):
    """Same constructor / forward contract as the reference model (model.py:130-276), inference regime."""

    default_normalization = {"mean": (0.485, 0.456, 0.406), "std": (0.229, 0.224, 0.225)}

    def __init__(self, logger=None, is_cuda=True, merge_branch_parameters=False, use_group_norm=False,
                 backbone_arch="resnet50", use_inverse_geom_model=True, simplify_affine=False, img_normalization=None):
        super(Os2dModel, self).__init__()
        if use_group_norm:
            raise NotImplementedError("group-norm backbones are outside the hot path and not provided")
        self.logger = logger or logging.getLogger("OS2D")
        self.img_normalization = img_normalization or self.default_normalization
        self.net_feature_maps = ResNetC4(backbone_arch)
        self.merge_branch_parameters = merge_branch_parameters
        extractor = self.net_feature_maps if merge_branch_parameters else ResNetC4(backbone_arch)
        self.simplify_affine = simplify_affine
        self.use_inverse_geom_model = use_inverse_geom_model
        self.os2d_head_creator = build_os2d_head_creator(simplify_affine, is_cuda, use_inverse_geom_model,
                                                         self.net_feature_maps.feature_map_stride,
                                                         self.net_feature_maps.feature_map_receptive_field)
        self.net_label_features = LabelFeatureExtractor(feature_extractor=extractor)
        self.eval()
        self.is_cuda = is_cuda
        if is_cuda:
            self.cuda()

    def apply_class_heads_to_feature_maps(self, feature_maps, class_head):
        """Flatten the head outputs over space (model.py:197-233)."""
        B = feature_maps.size(0)
        loc, cls, cls_detached, corners = class_head(feature_maps)
        C = cls.size(1)
        return (loc.reshape(B, C, 4, -1), cls.reshape(B, C, -1), cls_detached.reshape(B, C, -1),
                corners.reshape(B, C, 8, -1))

    def forward(self, images=None, class_images=None, feature_maps=None, class_head=None, train_mode=False,
                fine_tune_features=True):
        if train_mode:
            raise RuntimeError("os2d_b200.Os2dModel implements the evaluation regime only (train_mode=False)")
        with torch.no_grad():
            if feature_maps is None:
                assert images is not None, "If feature_maps is None than images cannot be None"
                if getattr(self, "use_packed_feature_maps", False):
                    feature_maps = self.net_feature_maps.forward_packed(images)
                else:
                    feature_maps = self.net_feature_maps(images)
            if class_head is None:
                assert class_images is not None, "If class_conv_layer is None than class_images cannot be None"
                class_head = self.os2d_head_creator.create_os2d_head(self.net_label_features(class_images))
            loc, cls, cls_detached, corners = self.apply_class_heads_to_feature_maps(feature_maps, class_head)
        fm_size = FeatureMapSize(w=feature_maps.size(3), h=feature_maps.size(2))
        return loc, cls, cls_detached, fm_size, corners

    def get_feature_map_size(self, img_size):
        """The reference runs a dummy image through the backbone (model.py:98-120, 278-288); for the C4 ResNets
        the result is ceil(size / 16) per side for every size >= 1."""
        return FeatureMapSize(w=int(math.ceil(img_size.w / 16)), h=int(math.ceil(img_size.h / 16)))
