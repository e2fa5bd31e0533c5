"""CPU, build container only: the oracle against the LIVE reference imported from /root/reference (skipped where the
reference tree does not exist, e.g. on the GPU box)."""
import sys
import warnings

import numpy as np
import pytest
import torch

from oracle import head_oracle as ho
from oracle import postproc_oracle as po
from _util import have_reference, rel_to_max, synth_inputs, REFERENCE_ROOT

pytestmark = pytest.mark.skipif(not have_reference(), reason="/root/reference not present")


def _ref_modules():
    warnings.filterwarnings("ignore")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from os2d.modeling.head import build_os2d_head_creator
    from os2d.structures.feature_map import FeatureMapSize
    return build_os2d_head_creator, FeatureMapSize


@pytest.mark.parametrize("simple,inverse", [(False, True), (True, False)])
def test_head_oracle_vs_live_reference(simple, inverse):
    build, FMS = _ref_modules()
    P = 4 if simple else 6
    tn = ho.random_transform_net(P, seed=3, spread=0.004)
    cms, fm = synth_inputs(5, 2, 9, 14, [(15, 15), (11, 21), (22, 10), (15, 16)], D=32)
    hc = build(simple, False, inverse, FMS(w=16, h=16), FMS(w=16, h=16))
    sd = dict(tn)
    sd["conv.1.num_batches_tracked"] = torch.tensor(0)
    sd["conv.4.num_batches_tracked"] = torch.tensor(0)
    hc.aligner.parameter_regressor.load_state_dict(sd)
    hc.eval()
    with torch.no_grad():
        head = hc.create_os2d_head(cms)
        loc, rec, rec2, corners = head(fm)
    cf = ho.prepare_class_features(cms)
    oloc, osc, ocor = ho.head_forward(cf, fm, tn, simple, inverse)
    assert rel_to_max(cf, head.class_feature_maps) < 1e-6
    assert rel_to_max(osc, rec) < 5e-6
    assert rel_to_max(oloc, loc) < 5e-5
    assert rel_to_max(ocor, corners) < 5e-6


def test_identity_init_gives_zero_loc():
    """Default TransformNet init regresses the identity (head.py:631-642): loc == 0, corners == anchor corners."""
    build, FMS = _ref_modules()
    hc = build(False, False, True, FMS(w=16, h=16), FMS(w=16, h=16))
    hc.eval()
    tn = {k: v.detach().clone() for k, v in hc.aligner.parameter_regressor.state_dict().items() if v.dtype.is_floating_point}
    cms, fm = synth_inputs(6, 1, 6, 7, [(15, 15)], D=16)
    cf = ho.prepare_class_features(cms)
    loc, score, corners = ho.head_forward(cf, fm, tn, False, True)
    assert float(loc.abs().max()) < 1e-5
    with torch.no_grad():
        rloc, rrec, _, _ = hc.create_os2d_head(cms)(fm)
    assert rel_to_max(score, rrec) < 5e-6


def test_c_nms_matches_torchvision():
    from torchvision.ops import nms as tv_nms
    g = torch.Generator().manual_seed(0)
    for n, quant in ((1, 0), (300, 0), (4000, 16), (2500, 4)):
        ctr = torch.rand(n, 2, generator=g) * 300
        wh = torch.rand(n, 2, generator=g) * 100 + 5
        boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
        scores = torch.rand(n, generator=g)
        if quant:
            scores = (scores * quant).round() / quant      # heavy ties
        ref = tv_nms(boxes, scores, 0.3).numpy()
        got = po.greedy_nms(boxes.numpy(), scores.numpy(), 0.3)
        np.testing.assert_array_equal(got, ref)
