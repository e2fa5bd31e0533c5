"""2-GPU probe: symmetric memory rendezvous, peer tensor views, copy-engine push bandwidth, device barrier, NCCL all-gather
timing.  Run under torchrun --nproc-per-node 2."""
import os, json
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
out = {"rank": rank, "world": world, "ndev": torch.cuda.device_count()}
try:
    out["peer_access"] = [torch.cuda.can_device_access_peer(lr, j) for j in range(torch.cuda.device_count()) if j != lr]
except Exception as e:
    out["peer_access_err"] = repr(e)
try:
    import torch.distributed._symmetric_memory as symm_mem
    n = 13 * 6400 * 100
    buf = symm_mem.empty(world, n, dtype=torch.float32, device=dev)
    buf.fill_(float(rank))
    h = symm_mem.rendezvous(buf, dist.group.WORLD)
    out["symm"] = {"ptrs": [hex(p) for p in h.buffer_ptrs], "has_mc": bool(getattr(h, "multicast_ptr", 0))}
    peers = [h.get_buffer(r, (world, n), torch.float32) for r in range(world)]
    h.barrier()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    # push my slice into every peer's buffer (copy engines)
    def push():
        for r in range(world):
            if r != rank:
                peers[r][rank].copy_(buf[rank], non_blocking=True)
    with torch.cuda.stream(st):
        for _ in range(3):
            push()
        h.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            push()
        e1.record()
        h.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out["ce_push_ms"] = ms
    out["ce_push_gbs"] = n * 4 * (world - 1) / ms / 1e6
    ok = all(bool((buf[r] == float(r)).all()) for r in range(world))
    out["ce_push_correct"] = ok
    # NCCL in-place all-gather of the same buffer
    flat = buf.view(-1)
    for _ in range(3):
        dist.all_gather_into_tensor(flat, buf[rank].reshape(-1))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        dist.all_gather_into_tensor(flat, buf[rank].reshape(-1))
    e1.record()
    torch.cuda.synchronize()
    out["nccl_allgather_ms"] = e0.elapsed_time(e1) / 10
except Exception as e:
    import traceback
    out["symm_err"] = traceback.format_exc()[-1500:]
print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
