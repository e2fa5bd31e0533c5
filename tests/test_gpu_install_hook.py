"""GPU: the UNMODIFIED reference `Os2dModel` (imported from /root/reference in the build container, from baseline/_ref -
tools/install_reference.sh - on the GPU box) carried by `os2d_b200.install.install()`: its forward(images, class_images) and
its box coder's decode_pyramid run through this package's kernels and agree with the same reference model un-hooked on the
CPU.  Runs in a subprocess because install() rebinds module globals."""
import os
import subprocess
import sys

import pytest

from _util import ROOT

pytestmark = pytest.mark.gpu


def _reference_root():
    for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "os2d", "modeling")):
            return cand
    return None


SCRIPT = r"""
import sys, warnings, logging, copy
warnings.filterwarnings("ignore")
sys.path.insert(0, %r); sys.path.insert(0, %r)
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import os2d.modeling.model as ref_model
from os2d.modeling.box_coder import Os2dBoxCoder
from os2d.structures.feature_map import FeatureMapSize
log = logging.getLogger("t")
torch.manual_seed(0)
# 1. the plain reference on the CPU (built BEFORE the hook rebinds the names)
plain = ref_model.Os2dModel(logger=log, is_cuda=False, backbone_arch="resnet50", use_inverse_geom_model=True, simplify_affine=False)
with torch.no_grad():
    lin = plain.os2d_head_creator.aligner.parameter_regressor.linear
    lin.weight.normal_(0, 0.004)                      # non-identity transforms
plain.eval()
sd = copy.deepcopy(plain.state_dict())
g = torch.Generator().manual_seed(1)
images = torch.randn(1, 3, 272, 336, generator=g)
class_images = [torch.randn(3, 96, 80, generator=g), torch.randn(3, 64, 128, generator=g), torch.randn(3, 80, 80, generator=g)]
with torch.no_grad():
    rloc, rcls, _, rsize, rcorners = plain(images, class_images)
plain_coder = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, plain.os2d_head_creator.box_grid_generator_image_level, plain.get_feature_map_size)
img_size = FeatureMapSize(w=336, h=272)
thr = float(rcls.median())
rdets = plain_coder.decode_pyramid([rloc[0]], [rcls[0]], [img_size], [5, 2, 5], nms_score_threshold=thr,
                                   nms_iou_threshold=0.3, transform_corners_pyramid=[rcorners[0]])

# 2. the same reference classes, hooked, on the GPU
import os2d_b200.install as hook
import os2d_b200.head as bh
hook.install()
net = ref_model.Os2dModel(logger=log, is_cuda=True, backbone_arch="resnet50", use_inverse_geom_model=True, simplify_affine=False)
assert isinstance(net.os2d_head_creator, bh.Os2dHeadCreator)
missing = net.load_state_dict(sd, strict=True)
net.eval()
with torch.no_grad():
    loc, cls, cls2, size, corners = net(images.cuda(), [c.cuda() for c in class_images])
assert cls2 is cls or torch.equal(cls2, cls)
assert size == rsize and loc.shape == rloc.shape and corners.shape == rcorners.shape


def rel(a, b):
    return float((a.cpu() - b).abs().max() / b.abs().max())


errs = {"score": rel(cls, rcls), "loc": rel(loc, rloc), "corners": rel(corners, rcorners)}
assert all(v < 1e-3 for v in errs.values()), errs
coder = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, net.os2d_head_creator.box_grid_generator_image_level, net.get_feature_map_size)
# decode the REFERENCE's scores on the GPU path: the selection must agree with the plain reference bit for bit
dets = coder.decode_pyramid([rloc[0].cuda()], [rcls[0].cuda()], [img_size], [5, 2, 5], nms_score_threshold=thr,
                            nms_iou_threshold=0.3, transform_corners_pyramid=[rcorners[0].cuda()])
assert len(dets) == len(rdets) and len(dets) > 0, (len(dets), len(rdets))
assert torch.equal(dets.get_field("scores").cpu(), rdets.get_field("scores"))
assert torch.equal(dets.get_field("labels").cpu(), rdets.get_field("labels"))
assert float((dets.bbox_xyxy.cpu() - rdets.bbox_xyxy).abs().max()) < 1e-2
print("HOOK_GPU_OK", errs, len(dets))
"""


@pytest.mark.skipif(_reference_root() is None, reason="the reference package is neither at /root/reference nor at baseline/_ref")
def test_reference_os2dmodel_runs_through_the_hook_on_gpu():
    out = subprocess.run([sys.executable, "-c", SCRIPT % (_reference_root(), ROOT)], capture_output=True, text=True, timeout=600)
    assert "HOOK_GPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
