"""Drop-in replacements of the reference head classes (os2d/modeling/head.py) whose arithmetic runs in
the sm_100a kernels behind the C ABI (include/os2d_b200.h):

    build_os2d_head_creator  head.py:12     TransformationNet  head.py:604
    Os2dAlignment            head.py:43     Os2dHeadCreator    head.py:204     Os2dHead  head.py:271

Constructor signatures, attribute names and state-dict keys
(``aligner.parameter_regressor.{conv.0,conv.1,conv.3,conv.4,linear}.*``) match the reference so reference
checkpoints load unchanged and ``Os2dModel`` can own an ``Os2dHeadCreator`` of this module.

Inference only: the kernels implement the eval-mode forward (BatchNorm running statistics, no autograd).
Calling the head with gradients required or BatchNorm in training mode raises - there is no PyTorch / CPU
fallback path in this package.
"""
import contextlib
import ctypes
import math
import os

import torch
import torch.nn as nn

from . import _cabi
from .structures import FeatureMapSize
from .box_coder import BoxGridGenerator

GRID = 15
CORR_CH = GRID * GRID
CORR_PAD = 240
Z_CHUNKS = CORR_PAD // 8
SCALE_Z = 64.0
SCALE_MEAN = 8.0
LO_SCALE = 2048.0
DC_CH = 225


class PackedFeatureMaps:
    """Image feature maps already in the head's operand layout: ``packed`` [B, H*W, D] fp16 = 32 * f / (||f|| + 1e-5)
    (head.py:339), written by the channels-last backbone tail (os2d_b200.model.ResNetC4.forward_packed,
    os2d_pack_image_features_nhwc).  ``Os2dHead.forward`` accepts it in place of the fp32 [B,D,H,W] tensor and skips its own
    pack kernel; ``size(i)`` / ``shape`` / ``device`` mirror the tensor the reference would pass."""

    def __init__(self, packed, height, width):
        assert packed.dim() == 3 and packed.dtype == torch.float16 and packed.size(1) == height * width
        self.packed = packed.contiguous()
        self.shape = torch.Size((packed.size(0), packed.size(2), height, width))
        self.device = packed.device
        self.requires_grad = False

    def size(self, i=None):
        return self.shape if i is None else self.shape[i]

    def __getitem__(self, idx):
        """Image sub-batch (the evaluation iterator slices the batch dimension)."""
        sub = self.packed[idx]
        if sub.dim() == 2:
            sub = sub.unsqueeze(0)
        return PackedFeatureMaps(sub, self.shape[2], self.shape[3])


def build_os2d_head_creator(do_simple_affine, is_cuda, use_inverse_geom_model, feature_map_stride,
                            feature_map_receptive_field):
    """Same factory as the reference (head.py:12-15)."""
    aligner = Os2dAlignment(do_simple_affine, is_cuda, use_inverse_geom_model)
    return Os2dHeadCreator(aligner, feature_map_stride, feature_map_receptive_field)


def _pow2_scale(t):
    """Power of two s such that max|t| * s lies in [0.5, 1) (exact rescale for fp16 storage)."""
    m = float(t.abs().max())
    if not math.isfinite(m) or m <= 0.0:
        return 1.0
    return 2.0 ** (-math.floor(math.log2(m)) - 1)


def _to_blob(wp, ks):
    """[128 rows, ci_pad (multiple of 16), ks, ks] fp32 -> shared-memory image consumed by csrc/conv.cu:
    [sub-chunk(16 ci)][dy][dx][k-group(8 ci)][row][8 ci] fp16."""
    rows, ci_pad = wp.shape[0], wp.shape[1]
    assert rows == 128 and ci_pad % 16 == 0
    v = wp.view(128, ci_pad // 16, 2, 8, ks, ks).permute(1, 4, 5, 2, 0, 3).contiguous()
    return v.to(torch.float16).contiguous()


class TransformationNet(nn.Module):
    """Parameter regression network, same modules / state-dict keys as the reference (head.py:604-661):
    conv = [Conv2d(225,128,7,p3), BN, ReLU, Conv2d(128,64,5,p2), BN, ReLU], linear = Conv2d(64,P,5,p2),
    last layer initialised to the identity transform (head.py:631-642)."""

    def __init__(self, output_dim=6, use_cuda=True, normalization='batchnorm', kernel_sizes=[7, 5], channels=[128, 64],
                 input_feature_dim=15 * 15, num_groups=16):
        super(TransformationNet, self).__init__()
        if normalization.lower() != 'batchnorm' or list(kernel_sizes) != [7, 5] or list(channels) != [128, 64] \
                or input_feature_dim != CORR_CH or output_dim not in (4, 6):
            raise NotImplementedError("os2d_b200 kernels implement the OS2D TransformNet (225->128 k7, 128->64 k5, "
                                      "64->{4,6} k5, batchnorm) only")
        mods = []
        ch_in = input_feature_dim
        for ch_out, k in zip(channels, kernel_sizes):
            mods += [nn.Conv2d(ch_in, ch_out, kernel_size=k, padding=k // 2), nn.BatchNorm2d(ch_out), nn.ReLU(inplace=True)]
            ch_in = ch_out
        self.conv = nn.Sequential(*mods)
        self.linear = nn.Conv2d(ch_in, output_dim, kernel_size=(kernel_sizes[-1], kernel_sizes[-1]),
                                padding=kernel_sizes[-1] // 2)
        with torch.no_grad():
            self.linear.weight.zero_()
            self.linear.bias.zero_()
            self.linear.bias[0] = 1
            self.linear.bias[4 if output_dim == 6 else 2] = 1
        self.output_dim = output_dim
        if use_cuda:
            self.conv.cuda()
            self.linear.cuda()
        self._packed = None
        self._packed_key = None

    def freeze_bn(self):
        for layer in self.modules():
            if isinstance(layer, nn.BatchNorm2d):
                layer.eval()

    # ---- weight folding / packing for the kernels (one-time per weight version) ----
    def _version_key(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def packed_weights(self):
        """BN-folded, fp16-packed operands of the three conv kernels (cached until a parameter changes)."""
        key = self._version_key()
        if self._packed is not None and self._packed_key == key:
            return self._packed
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d) and m.training:
                raise RuntimeError("os2d_b200: BatchNorm of the TransformNet is in training mode; the kernels implement "
                                   "the eval-mode forward only (call .eval() / freeze_bn())")
        self._packed = pack_transform_net(self.state_dict(), self.output_dim, self.linear.weight.device)
        self._packed_key = key
        return self._packed

    @_cabi.on_device_of
    def forward(self, corr_maps):
        """corr_maps [NB,225,H,W] fp32 -> transform parameters [NB,P,H,W] (head.py:648-655).  Stand-alone entry point
        (Os2dHead.forward never builds the fp32 correlation volume): the maps are normalised / packed by
        os2d_pack_corr_maps and run through the same three convolution kernels as the fused path."""
        _require_inference(corr_maps, self)
        NB, ch, H, W = corr_maps.shape
        assert ch == CORR_CH
        lib = _cabi.load()
        st = _cabi.stream_ptr()
        dev = corr_maps.device
        N = H * W
        corr = corr_maps.detach().to(torch.float32).contiguous()
        zvol = torch.empty(NB, Z_CHUNKS, N, 8, dtype=torch.float16, device=dev)
        rawvol = torch.empty(NB, CORR_CH, N, dtype=torch.float16, device=dev)
        _cabi.check(lib.os2d_pack_corr_maps(_cabi.ptr(corr), NB, H, W, _cabi.ptr(zvol), _cabi.ptr(rawvol), st),
                    "os2d_pack_corr_maps")
        return run_transform_convs(self.packed_weights(), self.output_dim, zvol, NB, H, W).view(NB, self.output_dim, H, W)


def _require_inference(t, module):
    if torch.is_grad_enabled() and (t.requires_grad or any(p.requires_grad for p in module.parameters())):
        raise RuntimeError("os2d_b200 implements inference only; call it under torch.no_grad() "
                           "(training / autograd through the head is out of scope and has no fallback)")
    if t.device.type != "cuda":
        raise RuntimeError("os2d_b200 requires CUDA tensors (no CPU path)")


_AUX_STREAMS = {}


_K3_STREAMS = {}


def _k3_stream(dev):
    """Side stream of the asynchronous resample kernel (Os2dHead.submit), one per device."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _K3_STREAMS:
        _K3_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _K3_STREAMS[key]


def _aux_stream(dev):
    """Side stream of the stage-concurrent correlation kernel, one per device."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _AUX_STREAMS:
        _AUX_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _AUX_STREAMS[key]


@_cabi.on_device_of
def run_transform_convs(pw, P, zvol, planes, H, W, timed=None, h1=None):
    """z volume [planes,30,N,8] -> parameters [planes,P,N] through the three conv kernels (os2d_transform_conv 1..3);
    with ``h1`` given, layer 1 has already run (os2d_correlate_conv1_concurrent)."""
    lib = _cabi.load()
    st = _cabi.stream_ptr()
    dev = zvol.device
    N = H * W
    call = timed if timed is not None else (lambda name, fn, *a: fn(*a))
    h2 = torch.empty(planes, 16, N, 8, dtype=torch.float16, device=dev)
    params = torch.empty(planes, P, N, dtype=torch.float32, device=dev)
    if h1 is None:
        h1 = torch.empty(planes, 16, N, 8, dtype=torch.float16, device=dev)
        _cabi.check(call("conv1", lib.os2d_transform_conv, 1, 128, _cabi.ptr(zvol), _cabi.ptr(pw["w1"]), _cabi.ptr(pw["alpha1"]),
                         _cabi.ptr(pw["beta1"]), _cabi.ptr(h1), planes, H, W, st), "os2d_transform_conv(1)")
    _cabi.check(call("conv2", lib.os2d_transform_conv, 2, 64, _cabi.ptr(h1), _cabi.ptr(pw["w2"]), _cabi.ptr(pw["alpha2"]),
                     _cabi.ptr(pw["beta2"]), _cabi.ptr(h2), planes, H, W, st), "os2d_transform_conv(2)")
    _cabi.check(call("conv3", lib.os2d_transform_conv, 3, P, _cabi.ptr(h2), _cabi.ptr(pw["w3"]), _cabi.ptr(pw["alpha3"]),
                     _cabi.ptr(pw["beta3"]), _cabi.ptr(params), planes, H, W, st), "os2d_transform_conv(3)")
    return params


def pack_transform_net(sd, out_dim, device):
    """Fold eval-mode BatchNorm into fp32 epilogue scale/shift and pack the conv weights for csrc/conv.cu.

    layer 1: rows = 128 output channels; input channels 0..224 = W1 * s1 (fp16); the K1 epilogue stores the
             *centred* z, so the per-pixel mean enters through DC side channels 225..227 carrying
             V[co,tap] = sum_ci W1 * s1 * 64 / 8 split into fp16 hi/lo (mean is split hi/lo as well).
    layer 2: rows 0..63 = fp16(W2 * s2), rows 64..127 = fp16 residual * 2048 (combined in the epilogue).
    layer 3: scatter-form operand of csrc/conv3s.cu: rows (tap, co), fp16 hi planes + unscaled fp16 residual planes.
    """
    f32 = dict((k, v.detach().to(device="cpu", dtype=torch.float64)) for k, v in sd.items() if v.dtype.is_floating_point)
    eps = 1e-5

    def bn_fold(conv, bn):
        a = f32[bn + ".weight"] / torch.sqrt(f32[bn + ".running_var"] + eps)
        b = (f32[conv + ".bias"] - f32[bn + ".running_mean"]) * a + f32[bn + ".bias"]
        return a, b

    out = {}
    # ---- layer 1 ----
    w1 = f32["conv.0.weight"]                      # [128,225,7,7]
    s1 = _pow2_scale(w1)
    wp = torch.zeros(128, CORR_PAD, 7, 7, dtype=torch.float64)
    wp[:, :CORR_CH] = w1 * s1
    v = w1.sum(dim=1) * (s1 * SCALE_Z / SCALE_MEAN)   # [128,7,7]
    vh = v.to(torch.float16).to(torch.float64)
    vl = (v - vh)
    wp[:, DC_CH] = vh
    wp[:, DC_CH + 1] = vl
    wp[:, DC_CH + 2] = vh
    a1, b1 = bn_fold("conv.0", "conv.1")
    out["w1"] = _to_blob(wp.float(), 7).to(device)
    out["alpha1"] = (a1 / (s1 * SCALE_Z)).float().to(device).contiguous()
    out["beta1"] = b1.float().to(device).contiguous()
    # ---- layer 2 ----
    w2 = f32["conv.3.weight"]                      # [64,128,5,5]
    s2 = _pow2_scale(w2)
    w2s = w2 * s2
    w2h = w2s.to(torch.float16).to(torch.float64)
    # accumulator row of (output channel c, part p): 32 * (c // 16) + 16 * p + c % 16, so that the hi and the residual
    # row of a channel sit in lanes l and l ^ 16 of the same epilogue warp (combined with one shuffle, csrc/conv.cu)
    cidx = torch.arange(64)
    row_hi = 32 * (cidx // 16) + (cidx % 16)
    row_lo = row_hi + 16
    wp = torch.zeros(128, 128, 5, 5, dtype=torch.float64)
    wp[row_hi] = w2h
    wp[row_lo] = (w2s - w2h) * LO_SCALE
    a2, b2 = bn_fold("conv.3", "conv.4")
    alpha2 = torch.zeros(128, dtype=torch.float64)
    beta2 = torch.zeros(128, dtype=torch.float64)
    alpha2[row_hi] = a2 / s2
    alpha2[row_lo] = a2 / s2
    beta2[row_hi] = b2
    beta2[row_lo] = b2
    out["w2"] = _to_blob(wp.float(), 5).to(device)
    out["alpha2"] = alpha2.float().to(device).contiguous()
    out["beta2"] = beta2.float().to(device).contiguous()
    out["row_hi2"] = row_hi
    # ---- layer 3 ----
    w3 = f32["linear.weight"]                      # [P,64,5,5]
    s3 = _pow2_scale(w3)
    w3s = w3 * s3
    w3h = w3s.to(torch.float16).to(torch.float64)
    w3l = (w3s - w3h).to(torch.float16).to(torch.float64)      # unscaled residual: |w3s| < 1 => abs error <= 2^-25
    # scatter-form operand (csrc/conv3s.cu): rows n = (dy*5 + dx) * P + co, [16 chunk8 (hi 0..7, lo 8..15)][NPAD][8]
    npad = (25 * out_dim + 15) // 16 * 16
    rows = torch.zeros(2, npad, 64, dtype=torch.float64)
    rows[0, :25 * out_dim] = w3h.permute(2, 3, 0, 1).reshape(25 * out_dim, 64)
    rows[1, :25 * out_dim] = w3l.permute(2, 3, 0, 1).reshape(25 * out_dim, 64)
    blob3 = rows.view(2, npad, 8, 8).permute(0, 2, 1, 3).reshape(16, npad, 8)
    a3 = torch.zeros(128, dtype=torch.float64)
    b3 = torch.zeros(128, dtype=torch.float64)
    a3[:] = 1.0 / s3
    b3[:out_dim] = f32["linear.bias"]
    out["w3"] = blob3.to(torch.float16).contiguous().to(device)
    out["alpha3"] = a3.float().to(device).contiguous()
    out["beta3"] = b3.float().to(device).contiguous()
    return out


class Os2dAlignment(nn.Module):
    """Transformation model: owns the parameter regressor and the transform conventions (head.py:43-193).
    The grid generation itself (prepare_transform_parameters_for_grid_sampler + F.affine_grid, head.py:81-193)
    is fused into the resample kernel (csrc/resample.cu)."""

    def __init__(self, do_simple_affine, is_cuda, use_inverse_geom_model):
        super(Os2dAlignment, self).__init__()
        self.model_type = "affine" if not do_simple_affine else "simple_affine"
        self.use_inverse_geom_model = use_inverse_geom_model
        transform_net_output_dim = 6 if self.model_type == "affine" else 4
        self.out_grid_size = FeatureMapSize(w=GRID, h=GRID)
        self.reference_feature_map_size = FeatureMapSize(w=GRID, h=GRID)
        self.network_stride = FeatureMapSize(w=1, h=1)
        self.network_receptive_field = FeatureMapSize(w=GRID, h=GRID)
        self.input_feature_dim = self.reference_feature_map_size.w * self.reference_feature_map_size.h
        self.parameter_regressor = TransformationNet(output_dim=transform_net_output_dim, use_cuda=is_cuda,
                                                     normalization='batchnorm', kernel_sizes=[7, 5], channels=[128, 64],
                                                     input_feature_dim=self.input_feature_dim)

    def prepare_transform_parameters_for_grid_sampler(self, transform_parameters):
        """[NB,P,H,W] -> [NB*H*W, 2, 3] affine matrices, inverted when use_inverse_geom_model (head.py:81-153; closed-form
        inverse instead of the batched LU).  Small tensor utility of the stand-alone API (elementwise torch ops); the
        fused path forms theta in registers (csrc/resample.cu)."""
        p = transform_parameters.permute(0, 2, 3, 1).reshape(-1, transform_parameters.size(1))
        z = torch.zeros_like(p[:, 0])
        if self.model_type == "affine":
            assert p.size(1) == 6, "Affine tranformation parameter vector has to be of dimension 6"
            a, b, tx, c, d, ty = (p[:, i] for i in range(6))
        else:
            assert p.size(1) == 4, "Simplified affine tranformation parameter vector has to be of dimension 4"
            a, b, tx, c, d, ty = p[:, 0], z, p[:, 1], z, p[:, 2], p[:, 3]
        if self.use_inverse_geom_model:
            det = a * d - b * c
            ia, ib, ic, id_ = d / det, -b / det, -c / det, a / det
            itx, ity = -(ia * tx + ib * ty), -(ic * tx + id_ * ty)
            singular = det == 0
            if bool(singular.any()):
                # failure handling of the reference (robust_inverse, head.py:123-134), per matrix like csrc/common.cuh
                e = 1e-5
                ar, dr, br, cr = a.double() + e, d.double() + e, b.double(), c.double()
                detr = ar * dr - br * cr
                ra, rb, rc, rd = dr / detr, -br / detr, -cr / detr, ar / detr
                rtx = -(ra * tx.double() + rb * ty.double()) / (1.0 + e)
                rty = -(rc * tx.double() + rd * ty.double()) / (1.0 + e)
                ia, ib = torch.where(singular, ra.float(), ia), torch.where(singular, rb.float(), ib)
                ic, id_ = torch.where(singular, rc.float(), ic), torch.where(singular, rd.float(), id_)
                itx, ity = torch.where(singular, rtx.float(), itx), torch.where(singular, rty.float(), ity)
            a, b, c, d, tx, ty = ia, ib, ic, id_, itx, ity
        return torch.stack([a, b, tx, c, d, ty], dim=1).view(-1, 2, 3)

    @_cabi.on_device_of
    def forward(self, corr_maps):
        """corr_maps [NB,225,H,W] -> grids of transformed points [NB,H,W,15,15,2] in the local coordinate system of every
        location (head.py:155-193).  Stand-alone entry point: the fused path never materialises this tensor."""
        params = self.parameter_regressor(corr_maps)
        NB, P, H, W = params.shape
        lib = _cabi.load()
        grids = torch.empty(NB, H, W, GRID, GRID, 2, dtype=torch.float32, device=params.device)
        params = params.contiguous()
        _cabi.check(lib.os2d_affine_grids(_cabi.ptr(params), NB, P, H, W, 1 if self.use_inverse_geom_model else 0,
                                          _cabi.ptr(grids), _cabi.stream_ptr()), "os2d_affine_grids")
        return grids


class Os2dHeadCreator(nn.Module):
    """Creates Os2dHead instances from class feature maps; owns the trainable aligner (head.py:204-268)."""

    def __init__(self, aligner, feature_map_stride, feature_map_receptive_field):
        super(Os2dHeadCreator, self).__init__()
        self.aligner = aligner
        rec_field, stride = self.get_rec_field_and_stride_after_concat_nets(
            feature_map_receptive_field, feature_map_stride, self.aligner.network_receptive_field, self.aligner.network_stride)
        self.box_grid_generator_image_level = BoxGridGenerator(box_size=rec_field, box_stride=stride)
        self.box_grid_generator_feature_map_level = BoxGridGenerator(box_size=self.aligner.network_receptive_field,
                                                                     box_stride=self.aligner.network_stride)

    @staticmethod
    def get_rec_field_and_stride_after_concat_nets(receptive_field_netA, stride_netA, receptive_field_netB, stride_netB):
        """Receptive field / stride of netB(netA(x)) (head.py:222-238)."""
        if hasattr(receptive_field_netA, "w"):
            rf_w, st_w = Os2dHeadCreator.get_rec_field_and_stride_after_concat_nets(
                receptive_field_netA.w, stride_netA.w, receptive_field_netB.w, stride_netB.w)
            rf_h, st_h = Os2dHeadCreator.get_rec_field_and_stride_after_concat_nets(
                receptive_field_netA.h, stride_netA.h, receptive_field_netB.h, stride_netB.h)
            return FeatureMapSize(w=rf_w, h=rf_h), FeatureMapSize(w=st_w, h=st_h)
        return stride_netA * (receptive_field_netB - 1) + receptive_field_netA, stride_netA * stride_netB

    @staticmethod
    def resize_feature_maps_to_reference_size(ref_size, feature_maps):
        """Bilinear resize of every class map to 15x15 (head.py:240-259); returns the un-normalised fp32 maps."""
        cf32, _ = _prepare_class_operands(feature_maps, normalized=False)
        return cf32

    def create_os2d_head(self, class_feature_maps):
        return Os2dHead(class_feature_maps, self.aligner, self.box_grid_generator_image_level,
                        self.box_grid_generator_feature_map_level, _from_raw_maps=True)


@_cabi.on_device_of
def _prepare_class_operands(feature_maps, normalized=True):
    """list of [1,D,h,w] CUDA fp32 maps -> (cf32 [C,D,15,15], packed fp16 [C,240,D]): ONE kernel launch for the whole
    (possibly ragged) set, os2d_pack_class_features_ragged; os2d_pack_class_features when the list is a sliced batch."""
    lib = _cabi.load()
    assert len(feature_maps) > 0
    dev = feature_maps[0].device
    if dev.type != "cuda":
        raise RuntimeError("os2d_b200 requires CUDA tensors (no CPU path)")
    D = feature_maps[0].size(1)
    C = len(feature_maps)
    cf32 = torch.empty(C, D, GRID, GRID, dtype=torch.float32, device=dev)
    packed = torch.empty(C, CORR_PAD, D, dtype=torch.float16, device=dev)
    sizes = set()
    maps = []
    for fm in feature_maps:
        assert fm.size(0) == 1, "Can process only batches of size 1, but have {0}".format(fm.size(0))
        assert fm.size(1) == D
        if fm.device != dev:
            raise RuntimeError("all class feature maps must live on the same CUDA device")
        m = fm.detach()
        if m.dtype != torch.float32 or not m.is_contiguous():
            m = m.to(dtype=torch.float32).contiguous()
        maps.append(m)
        sizes.add((m.size(2), m.size(3)))
    norm = 1 if normalized else 0
    if len(sizes) == 1 and all(maps[i].data_ptr() + maps[i].numel() * 4 == maps[i + 1].data_ptr() for i in range(C - 1)):
        # slices of one contiguous [C,D,h,w] batch: uniform entry point, no descriptor upload
        (h, w), = sizes
        rc = lib.os2d_pack_class_features(ctypes.c_void_p(maps[0].data_ptr()), C, D, h, w, norm, _cabi.ptr(cf32),
                                          _cabi.ptr(packed), _cabi.stream_ptr())
        _cabi.check(rc, "os2d_pack_class_features")
    else:
        # one launch over per-class pointers and sizes (the maps stay where the backbone wrote them)
        ptrs = torch.tensor([m.data_ptr() for m in maps], dtype=torch.int64).to(dev)
        hw = torch.tensor([[m.size(2), m.size(3)] for m in maps], dtype=torch.int32).to(dev)
        rc = lib.os2d_pack_class_features_ragged(_cabi.ptr(ptrs), _cabi.ptr(hw), C, D, norm, _cabi.ptr(cf32),
                                                 _cabi.ptr(packed), _cabi.stream_ptr())
        _cabi.check(rc, "os2d_pack_class_features_ragged")
        # the kernel reads the maps asynchronously: keep them alive until it has run on this stream
        for m in maps:
            m.record_stream(torch.cuda.current_stream())
    return cf32, packed


class Os2dHead(nn.Module):
    """Recognition + localisation scores of a batch of class feature maps against image feature maps
    (head.py:271-435).  ``forward`` has the reference signature and return tuple."""

    def __init__(self, class_feature_maps, aligner, box_grid_generator_image_level, box_grid_generator_feature_map_level,
                 pool_border_width=2, _from_raw_maps=False):
        super(Os2dHead, self).__init__()
        if pool_border_width != 2:
            raise NotImplementedError("the resample kernel pools the inner 11x11 grid points (pool_border_width=2)")
        if _from_raw_maps:
            maps = list(class_feature_maps)
        else:
            # reference constructor contract: an already resized [C,D,15,15] tensor (head.py:277-293)
            maps = [class_feature_maps[i:i + 1] for i in range(class_feature_maps.size(0))]
        cf32, packed = _prepare_class_operands(maps)
        self.class_feature_maps = cf32           # L2-normalised, as in head.py:293
        self._class_packed = packed
        self.class_batch_size = cf32.size(0)
        self.box_grid_generator_image_level = box_grid_generator_image_level
        self.box_grid_generator_feature_map_level = box_grid_generator_feature_map_level
        mask = torch.zeros(self.class_batch_size, 1, GRID, GRID, dtype=torch.float32, device=cf32.device)
        mask[:, :, pool_border_width:GRID - pool_border_width, pool_border_width:GRID - pool_border_width] = 1
        self.class_pool_mask = mask / mask.sum(dim=(2, 3), keepdim=True)     # head.py:295-302
        self.aligner = aligner
        self.max_planes_per_call = 4096
        self.concurrent_corr_sms = None  # see _concurrent_corr_sms
        self.workspace_bytes = None      # byte bound of the per-call workspace; None = half of the free device memory
        self._cmax_cache = {}
        self.supports_out_views = True
        # optional per-stage CUDA-event timing (bench.py): list of (stage, start_event, end_event) when not None
        self.profile_events = None

    def select(self, indices):
        """Head over a subset / reordering of this head's classes (shares the aligner; no re-packing of the operand).
        Used by the batched evaluation iterator when only some labels are searched in an image batch."""
        idx = torch.as_tensor(indices, dtype=torch.long, device=self.class_feature_maps.device)
        sub = Os2dHead.__new__(Os2dHead)
        nn.Module.__init__(sub)
        sub.class_feature_maps = self.class_feature_maps.index_select(0, idx)
        sub._class_packed = self._class_packed.index_select(0, idx).contiguous()
        sub.class_batch_size = int(idx.numel())
        sub.box_grid_generator_image_level = self.box_grid_generator_image_level
        sub.box_grid_generator_feature_map_level = self.box_grid_generator_feature_map_level
        sub.class_pool_mask = self.class_pool_mask.index_select(0, idx)
        sub.aligner = self.aligner
        sub.max_planes_per_call = self.max_planes_per_call
        sub.concurrent_corr_sms = self.concurrent_corr_sms
        sub.workspace_bytes = self.workspace_bytes
        sub._cmax_cache = {}
        sub.supports_out_views = True
        sub.profile_events = None
        return sub

    def _classes_per_chunk(self, B, N, P, dev):
        """Classes per workspace chunk: bounded in planes (max_planes_per_call) AND in bytes - one (image, class) plane of
        the z / raw / h1 / h2 / params volumes takes ~1.47 KB per location, so 1000+ class views on a 128x128..150x150
        pyramid level would otherwise ask for > 100 GB in one chunk (the reference's per-class loop never does)."""
        key = (B, N, P, self.max_planes_per_call, self.workspace_bytes)
        if key not in self._cmax_cache:
            per_plane = N * (Z_CHUNKS * 16 + CORR_CH * 2 + 256 + 256 + 4 * P)
            budget = self.workspace_bytes
            if budget is None:
                free, _ = torch.cuda.mem_get_info(dev)
                cached = torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
                budget = (free + cached) // 2
            self._cmax_cache[key] = max(1, min(self.max_planes_per_call, int(budget) // per_plane) // B)
        return self._cmax_cache[key]

    def _concurrent_corr_sms(self, planes, N):
        """SMs given to the correlation kernel when it runs NEXT TO conv1 (0 = one after the other).  ``concurrent_corr_sms``:
        None = environment OS2D_B200_CONCURRENT_CORR (default off), else the even SM count.  Only for problems with enough
        planes to keep both kernels busy."""
        n = self.concurrent_corr_sms
        if n is None:
            n = int(os.environ.get("OS2D_B200_CONCURRENT_CORR", "0") or 0)
        if n <= 0 or planes < 8 or planes * N < 40000:
            return 0
        return n & ~1

    def _timed(self, name, fn, *a):
        if self.profile_events is None:
            return fn(*a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*a)
        e1.record()
        self.profile_events.append((name, e0, e1))
        return rc

    def submit(self, feature_maps, out_views=None, _after_corr=None):
        """Pipelined form of ``forward`` for streams of images: identical results, but K3 (the L2-gather-bound resample / box
        kernel, no tensor work, no shared memory) is launched on a side stream, so its blocks run in the register / issue
        slack of the NEXT call's persistent tcgen05 kernels instead of after them.  Returns ``(outputs, event)``: outputs as
        ``forward`` (None with ``out_views``), valid once ``event`` has completed - e.g.
        ``torch.cuda.current_stream().wait_event(event)`` before consuming them."""
        return self.forward(feature_maps, out_views=out_views, _after_corr=_after_corr, _async_resample=True)

    @_cabi.on_device_of
    def forward(self, feature_maps, out_views=None, out_peers=None, _before_resample=None, _after_corr=None,
                _async_resample=False):
        """feature_maps [B,D,H,W] -> (loc [B,C,4,H,W], rec [B,C,1,H,W], rec_transform_detached (same tensor under
        no-grad, head.py:400-402), corners [B,C,8,H,W]).  ``out_views`` (extension, default None): (score, loc, corners)
        strided views [B,C,k,H*W] to write into instead of fresh tensors; the call then returns None.
        ``out_peers`` (extension, multi-GPU): ``(ptrs, slice_offset, per)`` - int64 CUDA tensor of the base pointers of every
        rank's [G,B,per,13,N] gather buffer (peer-mapped), the float offset of THIS rank's slice in such a buffer and the
        number of class slots per rank; K3 then stores its outputs into all those buffers itself (csrc/resample_p2p.cu)
        and the call returns None - the caller synchronises the ranks before reading; ``_before_resample`` (callable) is
        invoked once on the launch stream right before the first K3 launch (os2d_b200.dist enqueues the cross-rank
        "slot is free" barrier there, so K1 and the convolutions never wait for the other ranks)."""
        if torch.is_grad_enabled() and (feature_maps.requires_grad or
                                        any(p.requires_grad for p in self.aligner.parameters())):
            raise RuntimeError("os2d_b200.Os2dHead implements inference only; call it under torch.no_grad() "
                               "(training / autograd through the head is out of scope and has no fallback)")
        if feature_maps.device.type != "cuda":
            raise RuntimeError("os2d_b200 requires CUDA tensors (no CPU path)")
        B, D, H, W = feature_maps.shape
        assert D == self.class_feature_maps.size(1), \
            "Feature dimensionality of input={0} and class={1} feature maps has to equal".format(D, self.class_feature_maps.size(1))
        if H < 2 or W < 2:
            raise ValueError("feature map must be at least 2x2 (the reference divides by W-1, H-1: head.py:381-382)")
        C = self.class_batch_size
        dev = feature_maps.device
        lib = _cabi.load()
        st = _cabi.stream_ptr()
        N = H * W
        pw = self.aligner.parameter_regressor.packed_weights()
        P = self.aligner.parameter_regressor.output_dim
        inverse = 1 if self.aligner.use_inverse_geom_model else 0
        gen = self.box_grid_generator_image_level

        if isinstance(feature_maps, PackedFeatureMaps):
            img_packed = feature_maps.packed        # the producer already wrote the normalised fp16 operand
        else:
            fm = feature_maps.detach().to(dtype=torch.float32).contiguous()
            img_packed = torch.empty(B, N, D, dtype=torch.float16, device=dev)
            inv_ws = torch.empty(B, N, dtype=torch.float32, device=dev)
            _cabi.check(self._timed("pack_image", lib.os2d_pack_image_features, _cabi.ptr(fm), B, D, N, _cabi.ptr(inv_ws),
                                    _cabi.ptr(img_packed), st), "os2d_pack_image_features")

        if out_peers is not None:
            assert out_views is None, "out_views and out_peers are exclusive"
            loc = score = corners = o_score = o_loc = o_corners = None
        elif out_views is None:
            loc = torch.empty(B, C, 4, H, W, dtype=torch.float32, device=dev)
            score = torch.empty(B, C, 1, H, W, dtype=torch.float32, device=dev)
            corners = torch.empty(B, C, 8, H, W, dtype=torch.float32, device=dev)
            o_score, o_loc, o_corners = score.view(B, C, 1, N), loc.view(B, C, 4, N), corners.view(B, C, 8, N)
        else:
            # caller-owned strided views [B, C, k, N] (inner [k, N] contiguous), e.g. slices of a gather buffer
            o_score, o_loc, o_corners = out_views
            for v, k in ((o_score, 1), (o_loc, 4), (o_corners, 8)):
                assert v.shape == (B, C, k, N) and v.stride(3) == 1 and v.stride(2) == N and v.dtype == torch.float32
            loc = score = corners = None

        # class chunks bound the workspace (z / raw / hidden volumes); classes are independent in eval mode
        cmax = self._classes_per_chunk(B, N, P, dev)
        for c0 in range(0, C, cmax):
            cc = min(cmax, C - c0)
            planes = B * cc
            zvol = torch.empty(planes, Z_CHUNKS, N, 8, dtype=torch.float16, device=dev)
            rawvol = torch.empty(planes, CORR_CH, N, dtype=torch.float16, device=dev)
            cls = self._class_packed[c0:c0 + cc]
            corr_sms = self._concurrent_corr_sms(planes, N)
            if corr_sms:
                # K1 on `corr_sms` SMs next to conv1 on the others: conv1 consumes every plane as soon as K1 has released it
                h1 = torch.empty(planes, 16, N, 8, dtype=torch.float16, device=dev)
                flags = torch.empty(planes, dtype=torch.int32, device=dev)
                _cabi.check(self._timed("corr+conv1", lib.os2d_correlate_conv1_concurrent, _cabi.ptr(img_packed), _cabi.ptr(cls),
                                        B, cc, D, H, W, _cabi.ptr(zvol), _cabi.ptr(rawvol), _cabi.ptr(pw["w1"]),
                                        _cabi.ptr(pw["alpha1"]), _cabi.ptr(pw["beta1"]), _cabi.ptr(h1), _cabi.ptr(flags), corr_sms,
                                        st, ctypes.c_void_p(_aux_stream(dev).cuda_stream)), "os2d_correlate_conv1_concurrent")
                params = run_transform_convs(pw, P, zvol, planes, H, W, timed=self._timed, h1=h1)
            else:
                _cabi.check(self._timed("corr", lib.os2d_correlate, _cabi.ptr(img_packed), _cabi.ptr(cls), B, cc, D, H, W,
                                        _cabi.ptr(zvol), _cabi.ptr(rawvol), st), "os2d_correlate")
                params = run_transform_convs(pw, P, zvol, planes, H, W, timed=self._timed)
            if _after_corr is not None:          # hook of os2d_b200.dist: K1 (of the first chunk) has been launched
                _after_corr()
                _after_corr = None
            if _before_resample is not None:
                _before_resample()
                _before_resample = None
            if out_peers is not None:
                # K3 + collective in one kernel: outputs go to this rank's slice of every rank's gather buffer
                ptrs, slice_off, per = out_peers
                for b in range(B):
                    base = int(slice_off) + (b * int(per) + c0) * 13 * N
                    _cabi.check(self._timed("resample", lib.os2d_resample_boxes_p2p, _cabi.ptr(rawvol[b * cc:]),
                                            _cabi.ptr(params[b * cc:]), cc, P, H, W, inverse, float(gen.box_stride.w),
                                            float(gen.box_stride.h), float(gen.box_size.w), float(gen.box_size.h),
                                            _cabi.ptr(ptrs), int(ptrs.numel()), base, base + N, base + 5 * N, 13 * N, st),
                                "os2d_resample_boxes_p2p")
                continue
            # K3 writes straight into the final tensors (or into caller-provided views, e.g. this rank's slice of the
            # all-gather buffer): one launch when the planes of this chunk are contiguous in the output, else one per image
            if cc == C and out_views is None:
                launches = [(0, planes)]
            else:
                launches = [(b, cc) for b in range(B)]
            if _async_resample:
                # K3 on the side stream, after this chunk's conv3; the workspaces it reads must not be handed out again by
                # the allocator before it has run there
                main = torch.cuda.current_stream()
                k3s = _k3_stream(dev)
                ev_c3 = torch.cuda.Event()
                ev_c3.record(main)
                k3s.wait_event(ev_c3)
                for t in (rawvol, params) + ((score, loc, corners) if out_views is None else ()):
                    t.record_stream(k3s)             # (caller-owned views are the caller's to keep alive)
                k3_ctx, st_k3 = torch.cuda.stream(k3s), ctypes.c_void_p(k3s.cuda_stream)
            else:
                k3_ctx, st_k3 = contextlib.nullcontext(), st
            with k3_ctx:
                for b0, npl in launches:
                    _cabi.check(self._timed("resample", lib.os2d_resample_boxes, _cabi.ptr(rawvol[b0 * cc:]),
                                            _cabi.ptr(params[b0 * cc:]), npl, P, H, W, inverse, float(gen.box_stride.w),
                                            float(gen.box_stride.h), float(gen.box_size.w), float(gen.box_size.h),
                                            ctypes.c_void_p(o_score[b0, c0].data_ptr()), ctypes.c_void_p(o_loc[b0, c0].data_ptr()),
                                            ctypes.c_void_p(o_corners[b0, c0].data_ptr()), o_score.stride(1), o_loc.stride(1),
                                            o_corners.stride(1), st_k3), "os2d_resample_boxes")
        outputs = None if (out_views is not None or out_peers is not None) else (loc, score, score, corners)
        if _async_resample:
            done = torch.cuda.Event()
            done.record(_k3_stream(dev))
            return outputs, done
        return outputs

    @staticmethod
    @_cabi.on_device_of
    def resample_of_correlation_map_fast(corr_maps, resampling_grids_grid_coord, class_pool_mask):
        """corr_maps [B,C,225,H,W], grids [B,C,H,W,15,15,2] (unit coordinates of the feature map), mask [C,1,15,15] ->
        pooled matches [B,C,1,H,W] (head.py:439-520; `_simple` :523-594 computes the same quantity).  Stand-alone entry
        point with explicit tensors; Os2dHead.forward fuses grid generation and sampling instead."""
        if corr_maps.device.type != "cuda":
            raise RuntimeError("os2d_b200 requires CUDA tensors (no CPU path)")
        B, C, ch, H, W = corr_maps.shape
        assert ch == CORR_CH and tuple(resampling_grids_grid_coord.shape) == (B, C, H, W, GRID, GRID, 2), \
            "the number of channels in the correlation map should match the size of the resampling grid"
        lib = _cabi.load()
        corr = corr_maps.detach().to(torch.float32).contiguous()
        grids = resampling_grids_grid_coord.detach().to(torch.float32).contiguous()
        mask = class_pool_mask.detach().to(torch.float32).reshape(C, CORR_CH).contiguous()
        out = torch.empty(B, C, 1, H, W, dtype=torch.float32, device=corr.device)
        _cabi.check(lib.os2d_resample_with_grid(_cabi.ptr(corr), _cabi.ptr(grids), _cabi.ptr(mask), B * C, C, H, W,
                                                _cabi.ptr(out), _cabi.stream_ptr()), "os2d_resample_with_grid")
        return out

    resample_of_correlation_map_simple = resample_of_correlation_map_fast


def normalize_feature_map_L2(feature_maps, epsilon=1e-6):
    """x / (||x||_2 over dim 1 + eps) (head.py:597-601); tensor utility kept for API compatibility."""
    return feature_maps / (feature_maps.norm(dim=1, keepdim=True) + epsilon)
