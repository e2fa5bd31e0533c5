"""CPU, build container only: the monkey-patch hook that routes the unmodified reference through this package
(run in a subprocess: install() rebinds module globals)."""
import subprocess
import sys

import pytest

from _util import have_reference, REFERENCE_ROOT, ROOT

pytestmark = pytest.mark.skipif(not have_reference(), reason="/root/reference not present")

SCRIPT = r"""
import sys, warnings, logging
warnings.filterwarnings("ignore")
sys.path.insert(0, %r); sys.path.insert(0, %r)
import torch
import os2d_b200.install as hook
import os2d_b200.head as bh
hook.install()
import os2d.modeling.model as ref_model
import os2d.modeling.head as ref_head
from os2d.modeling.box_coder import Os2dBoxCoder
from os2d.structures.feature_map import FeatureMapSize
assert ref_model.build_os2d_head_creator is bh.build_os2d_head_creator
assert ref_head.Os2dHead is bh.Os2dHead
# the UNMODIFIED reference Os2dModel now owns the B200 head creator and keeps its state-dict layout
net = ref_model.Os2dModel(logger=logging.getLogger("t"), is_cuda=False, backbone_arch="resnet50",
                          use_inverse_geom_model=True, simplify_affine=False)
assert isinstance(net.os2d_head_creator, bh.Os2dHeadCreator)
keys = [k for k in net.state_dict().keys() if k.startswith("os2d_head_creator.")]
assert len(keys) == 16 and "os2d_head_creator.aligner.parameter_regressor.conv.0.weight" in keys, keys
assert isinstance(net.os2d_head_creator.box_grid_generator_image_level.box_size, FeatureMapSize)
bc = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, net.os2d_head_creator.box_grid_generator_image_level, net.get_feature_map_size)
assert bc._get_default_boxes(FeatureMapSize(w=64, h=48)).bbox_xyxy.shape == (12, 4)
try:
    bc.decode_pyramid([torch.zeros(1, 4, 12)], [torch.zeros(1, 12)], [FeatureMapSize(w=64, h=48)], [0])
    raise SystemExit("decode on CPU tensors must fail loudly")
except RuntimeError as e:
    assert "CUDA" in str(e)
import os2d.data.voc_eval as ref_voc, os2d_b200.voc_eval as bv
assert ref_voc.do_voc_evaluation is bv.do_voc_evaluation
print("INSTALL_OK")
"""


def test_install_hook_rebinds_reference_names():
    out = subprocess.run([sys.executable, "-c", SCRIPT % (REFERENCE_ROOT, ROOT)], capture_output=True, text=True, timeout=300)
    assert "INSTALL_OK" in out.stdout, out.stdout + out.stderr
