"""Small invocations of the round-2 kernels for compute-sanitizer (tools/gpu_sanitize.sh): fused decode + NMS (single-chunk and
multi-chunk with the global final sort), gather, channels-last pack, K1 next to conv1, K3 with the peer sink on one device."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import synth_inputs  # noqa: E402
from oracle import head_oracle as ho  # noqa: E402
from os2d_b200 import _cabi  # noqa: E402
from os2d_b200 import head as bh  # noqa: E402
from os2d_b200.box_coder import BoxGridGenerator, Os2dBoxCoder, make_resize_transform  # noqa: E402
from os2d_b200.structures import FeatureMapSize  # noqa: E402

g = torch.Generator().manual_seed(0)
gen = BoxGridGenerator(FeatureMapSize(w=240, h=240), FeatureMapSize(w=16, h=16))
coder = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, gen, lambda s: FeatureMapSize(w=-(-s.w // 16), h=-(-s.h // 16)))
# 1. single-chunk decode + NMS, duplicated ids, corners
sides = [12, 20]
loc = [(torch.randn(4, 4, s * s, generator=g)).cuda() for s in sides]
cls = [torch.rand(4, s * s, generator=g).cuda() for s in sides]
cor = [(torch.randn(4, 8, s * s, generator=g) * 50).cuda() for s in sides]
sizes = [FeatureMapSize(w=16 * s, h=16 * s) for s in sides]
inv = [make_resize_transform(FeatureMapSize(w=320, h=320)) for _ in sides]
d = coder.decode_pyramid(loc, cls, sizes, [3, 1, 3, 0], nms_score_threshold=0.2, inverse_box_transforms=inv, transform_corners_pyramid=cor)
print("decode single-chunk:", len(d))
# 2. multi-chunk (> 10000 candidates per label): chunk loop + global final sort
sides = [64, 80]
loc = [(torch.randn(2, 4, s * s, generator=g) * 1.2).cuda() for s in sides]
cls = [torch.rand(2, s * s, generator=g).cuda() for s in sides]
sizes = [FeatureMapSize(w=16 * s, h=16 * s) for s in sides]
inv = [make_resize_transform(FeatureMapSize(w=1280, h=1280)) for _ in sides]
d = coder.decode_pyramid(loc, cls, sizes, [5, 5], nms_score_threshold=0.1, nms_iou_threshold=0.9, inverse_box_transforms=inv)
print("decode multi-chunk:", len(d))
# 3. channels-last pack
lib = _cabi.load()
a = torch.randn(70, 1024, generator=g).cuda()
b = torch.randn(70, 1024, generator=g).cuda()
out = torch.empty(70, 1024, dtype=torch.float16, device="cuda")
_cabi.check(lib.os2d_pack_image_features_nhwc(_cabi.ptr(a), _cabi.ptr(b), 0, 1, 70, 1024, _cabi.ptr(out), _cabi.stream_ptr()), "nhwc")
# 4. head with K1 next to conv1, then K3 on the side stream
tn = ho.random_transform_net(6, seed=1, spread=0.005)
cms, fm = synth_inputs(1, 1, 64, 64, [(15, 15)] * 10)
hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
hc.eval()
with torch.no_grad():
    head = hc.create_os2d_head([c.cuda() for c in cms])
    ref = head(fm.cuda())
    head.concurrent_corr_sms = 16
    out = head(fm.cuda())
    head.concurrent_corr_sms = 0
    (out2, ev) = head.submit(fm.cuda())
    ev.synchronize()
torch.cuda.synchronize()
print("concurrent == sequential:", all(torch.equal(x, y) for x, y in zip(ref, out)), "submit == forward:", all(torch.equal(x, y) for x, y in zip(ref, out2)))
print("== stage r02 PASSED")
