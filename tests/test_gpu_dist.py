"""GPU, needs >= 2 devices (skipped on a 1-GPU box): the class-sharded head with each of its three gather mechanisms
(copy-engine pushes through symmetric memory, K3 with peer stores, NCCL all-gather) == the single-GPU head, bit for bit -
synchronous and pipelined; the sharded host->device upload; decode + NMS sharded by label == single GPU."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(rank, world, port):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))


def _head_creator(seed=5):
    from os2d_b200 import head as bh
    from os2d_b200.structures import FeatureMapSize
    from oracle import head_oracle as ho
    tn = ho.random_transform_net(6, seed=seed, spread=0.005)
    hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
    hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
    hc.eval()
    return hc


def _worker(rank, world, port, ret, mode):
    import torch.distributed as dist
    from os2d_b200 import dist as bd
    from _util import synth_inputs
    _setup(rank, world, port)
    cms, fm = synth_inputs(77, 2, 24, 21, [(15, 15), (12, 18), (19, 11), (15, 15), (9, 9)])
    _, fm2 = synth_inputs(78, 2, 24, 21, [(15, 15)])
    hc = _head_creator()
    ok = True
    with torch.no_grad():
        maps = [c.cuda() for c in cms]
        sharded = bd.ClassShardedHead(maps, hc.create_os2d_head, gather=mode.split("+")[0])
        sharded.async_resample = mode.endswith("+async_k3")
        full = hc.create_os2d_head(maps)
        refs = [full(f.cuda()) for f in (fm, fm2)]
        # synchronous API, twice (buffer reuse)
        for f, (rloc, rscore, _, rcorners) in list(zip((fm, fm2), refs)) * 2:
            loc, score, corners = sharded(f.cuda())
            ok &= bool(torch.equal(loc, rloc) and torch.equal(score, rscore) and torch.equal(corners, rcorners))
        # pipelined API: two submits in flight, results read afterwards (views of the ring slots)
        h0 = sharded.submit(fm.cuda())
        h1 = sharded.submit(fm2.cuda())
        for h, (rloc, rscore, _, rcorners) in zip((h0, h1), refs):
            loc, score, corners = h.wait()
            ok &= bool(torch.equal(loc, rloc) and torch.equal(score, rscore) and torch.equal(corners, rcorners))
        for i in range(6):                     # slot reuse under load: alternate inputs, check every result
            h = sharded.submit((fm if i % 2 == 0 else fm2).cuda())
            loc, score, corners = h.wait()
            rloc, rscore, _, rcorners = refs[i % 2]
            ok &= bool(torch.equal(loc, rloc) and torch.equal(score, rscore) and torch.equal(corners, rcorners))
        sharded.drain()
    torch.cuda.synchronize()
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def _spawn(fn, *args, world=2):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(fn, args=(world, port, ret) + args, nprocs=world, join=True)
    return ret


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["copy_engine", "copy_engine+async_k3", "fused", "nccl"])
def test_class_sharded_head_equals_single_gpu(mode):
    ret = _spawn(_worker, mode)
    assert ret[0] and ret[1]


def _upload_worker(rank, world, port, ret):
    import torch.distributed as dist
    from os2d_b200 import dist as bd
    _setup(rank, world, port)
    g = torch.Generator().manual_seed(3)
    hosts = [torch.randn(1, 1024, 13, 17, generator=g).pin_memory() for _ in range(3)]   # numel not divisible by 64 * world
    up = bd.ShardedUpload(hosts[0].shape, torch.float32, torch.device("cuda", rank))
    ok = up.bytes_per_rank() < hosts[0].numel() * 4
    for i in range(7):
        t, ev = up.upload(hosts[i % 3])
        torch.cuda.current_stream().wait_event(ev)
        ok &= bool(torch.equal(t.cpu(), hosts[i % 3]))
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_upload_reassembles_the_host_tensor():
    ret = _spawn(_upload_worker)
    assert ret[0] and ret[1]


def _detector_worker(rank, world, port, ret):
    import torch.distributed as dist
    from os2d_b200 import dist as bd
    from os2d_b200.box_coder import Os2dBoxCoder
    from os2d_b200.structures import FeatureMapSize
    from _util import synth_inputs
    _setup(rank, world, port)
    cms, fm = synth_inputs(91, 1, 22, 26, [(15, 15), (12, 18), (19, 11), (15, 15), (9, 9), (15, 15)])
    class_ids = [7, 3, 7, 5, 9, 3]                     # duplicated ids: all views of a label must land on one rank
    hc = _head_creator(seed=6)
    coder = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, hc.box_grid_generator_image_level,
                         lambda s: FeatureMapSize(w=-(-s.w // 16), h=-(-s.h // 16)))
    img = FeatureMapSize(w=26 * 16, h=22 * 16)
    with torch.no_grad():
        maps = [c.cuda() for c in cms]
        det = bd.ClassShardedDetector(maps, class_ids, hc.create_os2d_head, coder)
        full = hc.create_os2d_head(maps)
        loc, score, _, corners = full(fm.cuda())
        thr = float(score.median())
        got = det([fm.cuda()], [img], nms_score_threshold=thr, nms_iou_threshold=0.3)
        n = 22 * 26
        ref = coder.decode_pyramid([loc[0].view(6, 4, n)], [score[0].view(6, n)], [img], class_ids, nms_score_threshold=thr,
                                   nms_iou_threshold=0.3, transform_corners_pyramid=[corners[0].view(6, 8, n)])
    # same detections; label blocks may come in a different order (rank order vs set order): compare label by label
    why = []
    if len(got) != len(ref) or len(ref) == 0:
        why.append("counts %d vs %d" % (len(got), len(ref)))
    for lab in set(class_ids):
        a = got.get_field("labels") == lab
        b = ref.get_field("labels") == lab
        if int(a.sum()) != int(b.sum()):
            why.append("label %d: %d vs %d detections" % (lab, int(a.sum()), int(b.sum())))
            continue
        for name, x, y in (("boxes", got.bbox_xyxy, ref.bbox_xyxy), ("scores", got.get_field("scores"), ref.get_field("scores")),
                           ("corners", got.get_field("transform_corners"), ref.get_field("transform_corners")),
                           ("anchors", got.get_field("default_boxes").bbox_xyxy, ref.get_field("default_boxes").bbox_xyxy)):
            if not torch.equal(x[a], y[b]):
                why.append("label %d: %s differ (max %g)" % (lab, name, float((x[a] - y[b]).abs().max())))
    ok = "; ".join(why)
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_label_sharded_detector_equals_single_gpu():
    ret = _spawn(_detector_worker)
    assert ret[0] == "" and ret[1] == "", (ret[0], ret[1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_tensors_on_a_non_current_device():
    """tensors on cuda:1 while cuda:0 is current - the entry points switch to the tensors' device (stream, SM
    count, per-device shared-memory attribute) and give the same bits as on cuda:0."""
    from os2d_b200.box_coder import Os2dBoxCoder
    from os2d_b200.structures import FeatureMapSize
    from _util import synth_inputs
    torch.cuda.set_device(0)
    cms, fm = synth_inputs(55, 1, 20, 23, [(15, 15), (12, 18), (19, 11)])
    outs = []
    for d in (1, 0):
        dev = torch.device("cuda", d)
        hc = _head_creator().to(dev)
        with torch.no_grad():
            head = hc.create_os2d_head([c.to(dev) for c in cms])
            loc, score, _, corners = head(fm.to(dev))
            coder = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, hc.box_grid_generator_image_level,
                                 lambda s: FeatureMapSize(w=-(-s.w // 16), h=-(-s.h // 16)))
            dets = coder.decode_pyramid([loc[0].view(3, 4, -1)], [score[0].view(3, -1)], [FeatureMapSize(w=23 * 16, h=20 * 16)],
                                        [0, 1, 2], nms_score_threshold=0.5)
        assert loc.device == dev and torch.cuda.current_device() == 0
        outs.append((loc.cpu(), score.cpu(), corners.cpu(), dets.bbox_xyxy.cpu(), dets.get_field("scores").cpu()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
