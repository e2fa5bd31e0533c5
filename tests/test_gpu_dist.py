"""GPU, needs >= 2 devices (skipped on a 1-GPU box): class-sharded head over NCCL == single-GPU head, bit for bit."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret, fused=False):
    import torch.distributed as dist
    from os2d_b200 import head as bh, dist as bd
    from os2d_b200.structures import FeatureMapSize
    from oracle import head_oracle as ho
    from _util import synth_inputs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    tn = ho.random_transform_net(6, seed=5, spread=0.005)
    cms, fm = synth_inputs(77, 2, 24, 21, [(15, 15), (12, 18), (19, 11), (15, 15), (9, 9)])
    hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
    hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
    hc.eval()
    with torch.no_grad():
        maps = [c.cuda() for c in cms]
        sharded = bd.ClassShardedHead(maps, hc.create_os2d_head, fused_gather=fused)
        loc, score, corners = sharded(fm.cuda())
        if fused:                                   # second call: buffer reuse + the leading barrier
            loc, score, corners = sharded(fm.cuda())
        full = hc.create_os2d_head(maps)
        rloc, rscore, _, rcorners = full(fm.cuda())
    torch.cuda.synchronize()
    ret[rank] = bool(torch.equal(loc, rloc) and torch.equal(score, rscore) and torch.equal(corners, rcorners))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_class_sharded_head_equals_single_gpu():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2 or not os.environ.get("OS2D_B200_TEST_FUSED_GATHER"),
                    reason="needs 2 GPUs; the K3 + peer-store path is experimental (set OS2D_B200_TEST_FUSED_GATHER=1, run under timeout)")
def test_fused_gather_equals_single_gpu():
    """K3 storing into every rank's symmetric gather buffer (csrc/resample_p2p.cu) + device-side barrier == single GPU."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret, True), nprocs=2, join=True)
    assert ret[0] and ret[1]
