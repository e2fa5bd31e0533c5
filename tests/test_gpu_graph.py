"""GPU: CUDA-graph replay of a head call (os2d_b200.graphed.GraphedHead) gives the bits of the eager call."""
import pytest
import torch

from _util import synth_inputs
from oracle import head_oracle as ho

pytestmark = pytest.mark.gpu


def test_graphed_head_replays_eager_bits():
    from os2d_b200 import head as bh, GraphedHead
    from os2d_b200.structures import FeatureMapSize
    tn = ho.random_transform_net(6, seed=4, spread=0.005)
    cms, fm = synth_inputs(8, 1, 32, 32, [(15, 15), (12, 18), (19, 11), (15, 15)], D=1024)
    hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
    hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
    hc.eval()
    with torch.no_grad():
        head = hc.create_os2d_head([m.cuda() for m in cms])
        fm1 = fm.cuda()
        fm2 = (fm * 0.7 + 0.1).cuda()
        eager1 = [t.clone() for t in head(fm1)]
        eager2 = [t.clone() for t in head(fm2)]
        graphed = GraphedHead(head, fm1)
        for fm_d, eager in ((fm1, eager1), (fm2, eager2), (fm1, eager1)):
            out = graphed(fm_d)
            torch.cuda.synchronize()
            for a, b in zip(out, eager):
                assert torch.equal(a, b)
        with pytest.raises(AssertionError):
            graphed(fm1[:, :, :16])
