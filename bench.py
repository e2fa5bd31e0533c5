#!/usr/bin/env python
"""Benchmark of the OS2D head hot path (BASELINE.json metric: query-classes/sec at 1280 px input).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation (oracle port), rank 0

A step = one pass of the hot path over one batch of synthetic input: image feature map [B,1024,80,80] (1280 px
input through a stride-16 C4 backbone, which is outside the path) against C query classes, V2 head
(affine + inverse): image L2-norm/pack -> correlation -> TransformNet -> resample/pool -> loc/corners
(+ the all-gather of per-class outputs when N > 1).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "query-classes/sec at 1280px input"
UNIT = "classes/s"
D = 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--classes", type=int, default=100, help="query classes per GPU (weak scaling)")
    ap.add_argument("--size", type=int, default=1280, help="input image side in pixels")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--cpu-sample-classes", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(args, n_gpus):
    fm = -(-args.size // 16)
    return {
        "workload": "configs[1]: {0}px synthetic input ({1}x{1}x1024 C4 feature map), {2} query classes per GPU, "
                    "ResNet-C4 feature dim 1024, V2 inverse-geom affine head".format(args.size, fm, args.classes),
        "classes_per_gpu": args.classes, "global_classes": args.classes * n_gpus, "batch": args.batch,
        "feature_map": [fm, fm], "parallelism": "class-sharded x{}".format(n_gpus),
        "cache": "inputs larger than L2: every step streams ~{:.2f} GB of intermediate volumes per GPU".format(
            args.batch * args.classes * fm * fm * (480 + 450 + 256 + 128 + 24 + 52) / 1e9),
    }


def synth(args, device, seed):
    """Synthetic features of the config's shape (no dataset / checkpoint offline): ReLU'd Gaussians like C4 outputs."""
    g = torch.Generator().manual_seed(seed)
    fm = -(-args.size // 16)
    cms = (torch.randn(args.classes, D, 15, 15, generator=g) * 0.5 + 0.2).relu()
    fmap = (torch.randn(args.batch, D, fm, fm, generator=g) * 0.5 + 0.2).relu()
    return cms, fmap


# ---------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md): nvidia-smi during the timed region
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / power / throttle reasons through NVML every 20 ms while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.err = index, [], False, None
        self.thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.smax = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as e:   # noqa: BLE001
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0, int(get_reasons(self.h))))
            except Exception as e:   # noqa: BLE001
                self.err = repr(e)
                break
            time.sleep(0.002)

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: {}".format(self.err)]}
        sm = sorted(s[0] for s in self.samples)
        mask = 0
        for s in self.samples:
            mask |= s[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.smax, "reasons": sorted(v for k, v in self.REASONS.items() if mask & k),
                "samples": len(sm), "power_w_max": max(s[1] for s in self.samples)}


# ---------------------------------------------------------------------------------------------------------------
def seeded_transform_net(out_dim, seed=0, spread=0.02):
    """Seeded synthetic TransformNet weights with the reference's state-dict keys (conv.0/1/3/4, linear) and a non-identity
    output; both arms of the bench use this one generator (there are no checkpoints offline)."""
    import math
    g = torch.Generator().manual_seed(seed)
    tn = {}

    def conv(name, co, ci, k):
        bound = 1.0 / math.sqrt(ci * k * k)
        tn[name + ".weight"] = (torch.rand(co, ci, k, k, generator=g) * 2 - 1) * bound
        tn[name + ".bias"] = (torch.rand(co, generator=g) * 2 - 1) * bound

    def bn(name, c):
        tn[name + ".weight"] = 0.5 + torch.rand(c, generator=g)
        tn[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
        tn[name + ".running_mean"] = 0.05 * torch.randn(c, generator=g)
        tn[name + ".running_var"] = 0.01 + 0.05 * torch.rand(c, generator=g)

    conv("conv.0", 128, 225, 7)
    bn("conv.1", 128)
    conv("conv.3", 64, 128, 5)
    bn("conv.4", 64)
    tn["linear.weight"] = spread * torch.randn(out_dim, 64, 5, 5, generator=g)
    bias = torch.zeros(out_dim)
    bias[0] = 1
    bias[4 if out_dim == 6 else 2] = 1
    tn["linear.bias"] = bias
    return tn


# reference arm / cpu baseline: the oracle port of the reference's CPU implementation, all host threads
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(args, steps, warmup, sample_classes):
    """Times the oracle port of Os2dHead.forward on the host.  The thread count is the best of a short sweep
    (all hardware threads is rarely the fastest on a 2-socket host); `cores` reports the count actually used."""
    from oracle import head_oracle as ho
    g = torch.Generator().manual_seed(0)
    fm = -(-args.size // 16)
    cms = [(torch.randn(1, D, 15, 15, generator=g) * 0.5 + 0.2).relu() for _ in range(sample_classes)]
    fmap = (torch.randn(args.batch, D, fm, fm, generator=g) * 0.5 + 0.2).relu()
    tn = seeded_transform_net(6, seed=1, spread=0.005)
    cf = ho.prepare_class_features(cms)
    ncpu = os.cpu_count() or 1

    def run_once():
        ho.head_forward(cf, fmap, tn, False, True, class_chunk=sample_classes)

    best_t, best_n = None, ncpu
    with torch.no_grad():
        for n in sorted(set(x for x in (8, 16, 32, 64, ncpu // 2, ncpu) if 1 <= x <= ncpu)):
            torch.set_num_threads(n)
            run_once()
            t0 = time.perf_counter()
            run_once()
            dt = time.perf_counter() - t0
            if best_t is None or dt < best_t:
                best_t, best_n = dt, n
        torch.set_num_threads(best_n)
        for _ in range(warmup):
            run_once()
        t0 = time.perf_counter()
        for _ in range(steps):
            run_once()
        dt = time.perf_counter() - t0
    rate = args.batch * sample_classes * steps / dt
    sample = ("{} steps x {} classes x batch {} at {}x{} feature map (oracle port of Os2dHead.forward, torch CPU fp32, "
              "{} threads = best of a sweep up to {} hardware threads)").format(steps, sample_classes, args.batch, fm, fm,
                                                                              best_n, ncpu)
    return rate, dt / steps * 1e3, sample, best_n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_classes = max(1, min(args.cpu_sample_classes, args.classes))
    steps = max(1, args.steps)
    rate, ms, sample, used = cpu_reference_rate(args, steps, max(1, min(args.warmup, 2)), sample_classes)
    cfg = workload_config(args, args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from os2d_b200 import head as bh
    from os2d_b200 import dist as bd
    from os2d_b200.structures import FeatureMapSize

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node {}".format(args.gpus)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    fm_side = -(-args.size // 16)
    N = fm_side * fm_side
    B, C = args.batch, args.classes
    cms, fmap = synth(args, dev, seed=1234 + rank)
    tn = seeded_transform_net(6, seed=1, spread=0.005)
    hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
    hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
    hc.eval()
    with torch.no_grad():
        head = hc.create_os2d_head([cms[i:i + 1].to(dev) for i in range(C)])
    fm_host = fmap.pin_memory()
    fm_dev = fmap.to(dev)
    # N > 1: this rank's K3 writes straight into its slice of a [world,B,C,13,N] gather buffer; the in-place NCCL
    # all-gather of image i is asynchronous and overlaps the kernels of image i+1 (two buffers in flight)
    gathers = [bd.allocate_gather_buffer(B, C * world, N, world, dev) for _ in range(2)] if world > 1 else None
    pending = [None, None]
    out_free = [None, None]        # events: the D2H reader of buffer k is done (e2e loop only)

    def step(fm_d, i=0):
        with torch.no_grad():
            if world == 1:
                loc, score, _, corners = head(fm_d)
                return loc, score, corners
            k = i % 2
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None
            if out_free[k] is not None:
                torch.cuda.current_stream().wait_event(out_free[k])
            s_v, l_v, c_v = bd.local_views(gathers[k], rank)
            head(fm_d, out_views=(s_v, l_v, c_v))
            pending[k] = bd.all_gather_outputs(gathers[k], async_op=True)
            return s_v, l_v, c_v

    def drain():
        for k in range(2):
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident throughput (value) ----
    for i in range(args.warmup):
        step(fm_dev, i)
    drain()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    head.profile_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(fm_dev, i)
    drain()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    stage_ms = {}
    for name, a, b in head.profile_events:
        stage_ms.setdefault(name, []).append(a.elapsed_time(b))
    head.profile_events = None
    stage_avg = {k: sum(v) / len(v) for k, v in stage_ms.items()}
    value = B * C * world * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public API with host buffers (e2e) ----
    # every step: H2D of the feature map from pinned memory, head (+ all-gather), D2H of this rank's loc/score/corners
    out_host = [torch.empty(B, C, 4, N).pin_memory(), torch.empty(B, C, 1, N).pin_memory(), torch.empty(B, C, 8, N).pin_memory()]
    s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
    s_main = torch.cuda.current_stream()
    fm_bufs = [torch.empty_like(fm_dev) for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]

    def e2e_steps(n):
        for i in range(n):
            k = i % 2
            with torch.cuda.stream(s_h2d):
                if i >= 2:
                    s_h2d.wait_event(ev_free[k])
                fm_bufs[k].copy_(fm_host, non_blocking=True)
                ev_in[k].record(s_h2d)
            s_main.wait_event(ev_in[k])
            outs = step(fm_bufs[k], i)
            ev_free[k].record(s_main)
            done = torch.cuda.Event()
            done.record(s_main)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(done)
                if world == 1:
                    loc, score, corners = outs
                    srcs = (loc.view(B, C, 4, N), score.view(B, C, 1, N), corners.view(B, C, 8, N))
                else:
                    srcs = (outs[1], outs[0], outs[2])
                for dst, src in zip(out_host, srcs):
                    dst.copy_(src, non_blocking=True)
                    if world == 1:
                        src.record_stream(s_d2h)
                if world > 1:
                    out_free[k] = torch.cuda.Event()
                    out_free[k].record(s_d2h)
        drain()
        s_main.wait_stream(s_d2h)
        out_free[0] = out_free[1] = None

    e2e_steps(max(2, args.warmup))
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_steps(args.steps)
    f1.record()
    barrier()
    e2e_ms = max_over_ranks(f0.elapsed_time(f1))
    e2e_value = B * C * world * args.steps / (e2e_ms * 1e-3)
    h2d_bytes = fm_host.numel() * 4
    d2h_bytes = sum(t.numel() for t in out_host) * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- second figure of the metric: decode + per-class NMS of one image's outputs (all anchors, threshold -inf) ----
    postproc = None
    try:
        from os2d_b200.box_coder import Os2dBoxCoder
        coder = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, hc.box_grid_generator_image_level,
                             lambda s: FeatureMapSize(w=-(-s.w // 16), h=-(-s.h // 16)))
        with torch.no_grad():
            drain()
            _l, _s, _, _c = head(fm_dev)
            loc, score, corners = _l, _s, _c
        img = FeatureMapSize(w=fm_side * 16, h=fm_side * 16)
        args_pp = ([loc[0].reshape(C, 4, N)], [score[0].reshape(C, N)], [img], list(range(C)))
        for _ in range(2):
            dets = coder.decode_pyramid(*args_pp, nms_score_threshold=float("-inf"), nms_iou_threshold=0.3,
                                        transform_corners_pyramid=[corners[0].reshape(C, 8, N)])
        torch.cuda.synchronize()
        p0 = time.perf_counter()
        for _ in range(5):
            dets = coder.decode_pyramid(*args_pp, nms_score_threshold=float("-inf"), nms_iou_threshold=0.3,
                                        transform_corners_pyramid=[corners[0].reshape(C, 8, N)])
        torch.cuda.synchronize()
        pp_ms = (time.perf_counter() - p0) / 5 * 1e3
        postproc = {"decode_nms_ms_per_image": pp_ms, "candidates": C * N, "detections": len(dets),
                    "classes_per_s_head_plus_postproc": C / ((ms_total / args.steps + pp_ms) * 1e-3),
                    "note": "rank-0 classes only, score threshold -inf (every anchor is an NMS candidate)"}
    except Exception as e:   # noqa: BLE001
        postproc = {"error": repr(e)}

    # ---- roofline of the dominant kernel (conv1 of the TransformNet) + the correlation GEMM ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside the step)" if peaks else "fallback 1.4 PF sustained"
    planes = B * C
    flop_conv1 = 2.0 * N * 225 * 128 * 49 * planes
    flop_corr = 2.0 * N * 225 * D * planes
    flop_conv2 = 2.0 * N * 128 * 64 * 25 * planes
    flop_conv3 = 2.0 * N * 64 * 6 * 25 * planes

    # DRAM traffic per launch from the committed ncu --set full capture (profiles/ncu_traffic.json, cfg-2 sized launch);
    # only reported when this run has the same launch geometry
    traffic = {}
    try:
        if B == 1 and C == 100 and fm_side == 80:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:   # noqa: BLE001
        traffic = {}
    kernel_names = {"conv1": "conv::conv_kernel<7>", "conv2": "conv::conv_kernel<5>", "corr": "corr::corr_kernel<1>"}

    def roof(name, flop):
        ms = stage_avg.get(name)
        if not ms:
            return None
        ach = flop / (ms * 1e-3) / 1e12
        tr = traffic.get(kernel_names.get(name, ""), {}).get("traffic_bytes")
        return {"kernel": kernel_names.get(name, name), "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": ach / peak_tf, "traffic": tr, "traffic_unit": "bytes per launch (ncu dram read+write)",
                "algorithmic_flop_per_launch": flop, "ms_per_launch": ms, "peak_source": peak_src}

    roofline = roof("conv1", flop_conv1)
    extra = {"roofline_corr": roof("corr", flop_corr), "roofline_conv2": roof("conv2", flop_conv2),
             "roofline_conv3": roof("conv3", flop_conv3),
             "stage_ms": stage_avg,
             "tensor_frac_whole_step": (flop_conv1 + flop_corr + flop_conv2 + flop_conv3) * args.steps / (ms_total * 1e-3) / 1e12 / peak_tf}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        sc = max(1, min(args.cpu_sample_classes, C))
        rate, _, sample, used = cpu_reference_rate(args, 3, 1, sc)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": used, "kind": "port", "sample": sample}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 operands, fp32 accumulate (hi/lo-split weights)", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": 6 * args.steps, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "postproc": postproc}
    line.update(extra)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
