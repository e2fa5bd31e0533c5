#!/usr/bin/env python
"""Benchmark of the OS2D head hot path (BASELINE.json metric: query-classes/sec at 1280 px input).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation on the host cores
    python bench.py --impl reference-gpu ...                 # (extra) the unfused reference on the GPU: cuBLAS/cuDNN path

A step = one pass of the hot path over one batch of synthetic input: image feature map [B,1024,80,80] (1280 px
input through a stride-16 C4 backbone, which is outside the path) against C query classes, V2 head
(affine + inverse): image L2-norm/pack -> correlation -> TransformNet -> resample/pool -> loc/corners
(+ the gather of per-class outputs when N > 1, through os2d_b200.dist.ClassShardedHead.submit - the public API).
Prints ONE JSON line on rank 0.  Sub-objects next to the headline (weak scaling, configs[1] per GPU):
  strong_c1000  configs[2]: 1000 classes sharded ceil(1000/N) per GPU (strong scaling)
  pipeline      head + decode + per-label NMS (+ gather of the survivors): the metric's second figure
  sustained     the same step looped for >= 3 s with its own clocks record
  parity        N > 1: every rank recomputes a foreign class block and compares it bit for bit with the gathered slice
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
ASYNC_RESAMPLE_DEFAULT = False
FM_SEED = 999          # the image feature map is replicated: same seed on every rank
CLASS_SEED = 1234      # + rank (weak) / 4321 + owner rank (strong)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--classes", type=int, default=100, help="query classes per GPU (weak scaling)")
    ap.add_argument("--size", type=int, default=1280, help="input image side in pixels")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--gather", default="copy_engine", choices=["copy_engine", "fused", "nccl"])
    ap.add_argument("--strong-classes", type=int, default=1000, help="global classes of the strong-scaling block (0 = skip)")
    ap.add_argument("--sustained-seconds", type=float, default=3.0, help="0 = skip the sustained block")
    ap.add_argument("--no-pipeline", action="store_true")
    ap.add_argument("--wave-classes", type=int, default=0,
                    help="run the five kernels class wave by class wave (this many classes per wave; 0 = all classes per kernel)")
    ap.add_argument("--concurrent-corr", type=int, default=-1,
                    help="SMs of the correlation kernel when it runs next to conv1 (0 = sequential kernels; -1 = library default)")
    ap.add_argument("--async-resample", type=int, default=-1,
                    help="1: K3 on a side stream (Os2dHead.submit) so that it overlaps the next image's tensor kernels; 0: in line")
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


def synth_fm(args):
    g = torch.Generator().manual_seed(FM_SEED)
    fm = -(-args.size // 16)
    return (torch.randn(args.batch, D, fm, fm, generator=g) * 0.5 + 0.2).relu()


def synth_classes(n, seed):
    """Synthetic class features (no dataset / checkpoint offline): ReLU'd Gaussians like C4 outputs."""
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n, D, 15, 15, generator=g) * 0.5 + 0.2).relu()


# ---------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md): NVML during the timed region
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / power / throttle reasons through NVML every few ms while the timed region runs."""
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
        return self

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
    output; every arm of the bench uses this one generator (there are no checkpoints offline)."""
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


# ---------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own Os2dHead.forward (UNMODIFIED os2d package from /root/reference or baseline/_ref) when it
# is importable, else the oracle port; all host threads it can use, a bounded sample of the workload per step
# ---------------------------------------------------------------------------------------------------------------
def import_reference():
    """The unmodified reference package: /root/reference in the build container, baseline/_ref (tools/install_reference.sh,
    travels with the gpurun snapshot) on the GPU box.  Returns the module of os2d.modeling.head or None."""
    import importlib
    import warnings
    for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "os2d", "modeling")):
            if cand not in sys.path:
                sys.path.insert(0, cand)
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    return importlib.import_module("os2d.modeling.head"), cand
            except Exception:   # noqa: BLE001
                if cand in sys.path:
                    sys.path.remove(cand)
    return None, None


def build_reference_head(ref_head_mod, cms, tn, is_cuda):
    from os2d.structures.feature_map import FeatureMapSize as RefSize
    hc = ref_head_mod.build_os2d_head_creator(False, is_cuda, True, RefSize(w=16, h=16), RefSize(w=16, h=16))
    sd = dict(tn)
    sd["conv.1.num_batches_tracked"] = torch.tensor(0)
    sd["conv.4.num_batches_tracked"] = torch.tensor(0)
    hc.aligner.parameter_regressor.load_state_dict(sd)
    hc.eval()
    if is_cuda:
        hc = hc.cuda()
    with torch.no_grad():
        return hc.create_os2d_head(cms)


def cpu_reference_rate(args, steps, warmup, sample_classes):
    """Times the reference's CPU implementation of Os2dHead.forward on the host.  The thread count is the best of a short
    sweep (all hardware threads is rarely the fastest); `cores` reports the count actually used."""
    import warnings
    fmap = synth_fm(args)
    fm = fmap.shape[-1]
    cls = synth_classes(sample_classes, CLASS_SEED)
    cms = [cls[i:i + 1] for i in range(sample_classes)]
    tn = seeded_transform_net(6, seed=1, spread=0.005)
    ncpu = os.cpu_count() or 1
    ref_mod, ref_from = import_reference()
    if ref_mod is not None:
        kind = "reference"
        head = build_reference_head(ref_mod, cms, tn, False)

        def run_once():
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                head(fmap)
        what = "unmodified os2d.modeling.head.Os2dHead.forward from {}".format(ref_from)
    else:
        kind = "port"
        from oracle import head_oracle as ho
        cf = ho.prepare_class_features(cms)

        def run_once():
            ho.head_forward(cf, fmap, tn, False, True, class_chunk=sample_classes)
        what = "oracle port of Os2dHead.forward"

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
    sample = ("{} steps x {} classes x batch {} at {}x{} feature map ({}, torch CPU fp32, {} threads = best of a sweep up to "
              "{} hardware threads)").format(steps, sample_classes, args.batch, fm, fm, what, best_n, ncpu)
    return rate, dt / steps * 1e3, sample, best_n, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_classes = max(1, min(args.cpu_sample_classes, args.classes))
    steps = max(1, args.steps)
    rate, ms, sample, used, kind = cpu_reference_rate(args, steps, max(1, min(args.warmup, 2)), sample_classes)
    cfg = workload_config(args, args.gpus)
    # the arm times a bounded sample: say so in the config it reports (the metric is a per-class rate)
    cfg["classes_timed_per_step"] = sample_classes
    cfg["same_config"] = sample_classes == args.classes
    cfg["note"] = "CPU arm: {} of the {} classes per step (bounded sample, rate is per class); one CPU run, not x{} GPUs".format(
        sample_classes, args.classes, args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_reference_gpu(args):
    """Extra arm (not part of the driver contract): the UNFUSED reference on the GPU - os2d's own Os2dHead.forward with
    is_cuda=True (cuBLAS bmm + cuDNN convolutions + grid_sample, head.py:342-435), the second baseline of BASELINE.md."""
    import warnings
    ref_mod, ref_from = import_reference()
    if ref_mod is None:
        print(json.dumps({"impl": "reference-gpu", "unavailable": "os2d reference not importable (no /root/reference, no baseline/_ref)"}))
        return
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    fmap = synth_fm(args).to(dev)
    C = args.classes
    cls = synth_classes(C, CLASS_SEED).to(dev)
    tn = seeded_transform_net(6, seed=1, spread=0.005)
    flags = {"cuda.matmul.allow_tf32": torch.backends.cuda.matmul.allow_tf32, "cudnn.allow_tf32": torch.backends.cudnn.allow_tf32,
             "cudnn.benchmark": True}
    torch.backends.cudnn.benchmark = True          # main.py:72
    head = build_reference_head(ref_mod, [cls[i:i + 1] for i in range(C)], tn, True)
    chunk = C
    def run_once():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with torch.no_grad():
                return head(fmap)
    # the reference materialises [C,225,H,W] fp32/fp64 volumes and [C,H,W,15,15,2] grids: find the largest class chunk that fits
    while True:
        try:
            sub = head if chunk == C else build_reference_head(ref_mod, [cls[i:i + 1] for i in range(chunk)], tn, True)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                with torch.no_grad():
                    sub(fmap)
            torch.cuda.synchronize()
            break
        except torch.cuda.OutOfMemoryError:
            torch.cuda.empty_cache()
            chunk = max(1, chunk // 2)
    heads = [build_reference_head(ref_mod, [cls[i:i + 1] for i in range(c0, min(C, c0 + chunk))], tn, True)
             for c0 in range(0, C, chunk)] if chunk != C else [head]

    def step():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with torch.no_grad():
                for h in heads:
                    h(fmap)
    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0).start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    cfg = workload_config(args, 1)
    cfg["classes_per_reference_call"] = chunk
    print(json.dumps({"impl": "reference-gpu", "metric": METRIC, "value": args.batch * C / (ms * 1e-3), "unit": UNIT, "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "dtype": "fp32 (bmm) / TF32 (cuDNN convs, torch default) / fp64 (grid_sample)", "data": "synthetic",
                      "config": cfg, "flags": flags, "clocks": sampler.stop(), "reference_from": ref_from,
                      "note": "unfused reference on the GPU: library kernels only, none of this repo's code"}), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from os2d_b200 import _cabi
    from os2d_b200 import head as bh
    from os2d_b200 import dist as bd
    from os2d_b200.structures import FeatureMapSize
    from os2d_b200.box_coder import Os2dBoxCoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node {}".format(args.gpus)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()

    fm_side = -(-args.size // 16)
    N = fm_side * fm_side
    B = args.batch
    fmap = synth_fm(args)
    fm_host = fmap.pin_memory()
    fm_dev = fmap.to(dev)
    tn = seeded_transform_net(6, seed=1, spread=0.005)
    hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
    hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
    hc.eval()

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

    def all_true(flag):
        if world > 1:
            t = torch.tensor([1 if flag else 0], device=dev, dtype=torch.int32)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return bool(t.item())
        return bool(flag)

    class Workload:
        """C_global classes sharded in contiguous blocks of `per` over the ranks; class block of rank r is seeded by
        (seed0 + r) so that any rank can regenerate a foreign block for the parity check."""

        def __init__(self, per, seed0, gather):
            self.per, self.seed0 = per, seed0
            self.C_global = per * world
            own = synth_classes(per, seed0 + rank).to(dev)
            with torch.no_grad():
                if world == 1:
                    self.head = hc.create_os2d_head([own[i:i + 1] for i in range(per)])
                    self.sharded = None
                else:
                    maps = [None] * self.C_global
                    for i in range(per):
                        maps[rank * per + i] = own[i:i + 1]
                    self.sharded = bd.ClassShardedHead(maps, hc.create_os2d_head, gather=gather)
                    self.head = self.sharded.head
            if args.wave_classes > 0:
                self.head.max_planes_per_call = args.wave_classes * B
            if args.concurrent_corr >= 0:
                self.head.concurrent_corr_sms = args.concurrent_corr
            self.async_k3 = (args.async_resample == 1) if args.async_resample >= 0 else ASYNC_RESAMPLE_DEFAULT
            if self.sharded is not None:
                self.sharded.async_resample = self.async_k3
            self.k3_events = []
            self.last_event = None

        def step(self, fm_d):
            """One pass of the hot path; returns the (score, loc, corners) views [B,per,k,N] of THIS rank's block."""
            with torch.no_grad():
                if self.sharded is None:
                    if self.async_k3:
                        (loc, score, _, corners), ev = self.head.submit(fm_d)
                        self.k3_events = self.k3_events[-3:] + [ev]
                        self.last_event = ev
                    else:
                        loc, score, _, corners = self.head(fm_d)
                        self.last_event = None
                    return score.view(B, self.per, 1, N), loc.view(B, self.per, 4, N), corners.view(B, self.per, 8, N)
                h = self.sharded.submit(fm_d)
                self.last = h
                self.last_event = h.local_event
                return h.local_views()

        def drain(self):
            if self.sharded is not None:
                self.sharded.drain()
            for ev in self.k3_events:
                torch.cuda.current_stream().wait_event(ev)
            self.k3_events = []
            if getattr(self, "last_event", None) is not None:
                torch.cuda.current_stream().wait_event(self.last_event)

        def timed(self, steps, warmup, with_stages=False, sampler=None):
            for _ in range(warmup):
                self.step(fm_dev)
            self.drain()
            barrier()
            if sampler is not None:
                sampler.start()
            if with_stages:
                self.head.profile_events = []
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = lib.os2d_b200_launch_count()
            barrier()
            e0.record()
            for _ in range(steps):
                self.step(fm_dev)
            self.drain()
            e1.record()
            barrier()
            launches = lib.os2d_b200_launch_count() - l0
            ms_total = max_over_ranks(e0.elapsed_time(e1))
            stage_avg = {}
            if with_stages:
                acc = {}
                for name, a, b in self.head.profile_events:
                    acc.setdefault(name, []).append(a.elapsed_time(b))
                self.head.profile_events = None
                stage_avg = {k: sum(v) / len(v) for k, v in acc.items()}
            return ms_total, stage_avg, launches

        def e2e(self, steps, warmup):
            """End to end through the public API with HOST buffers: every step uploads the feature map from pinned host
            memory (N > 1: each rank uploads 1/N of it over its own PCIe link, the parts are exchanged over NVLink -
            os2d_b200.dist.ShardedUpload), runs the head (+ gather) and downloads this rank's loc / score / corners."""
            per = self.per
            out_host = [torch.empty(B, per, 1, N).pin_memory(), torch.empty(B, per, 4, N).pin_memory(),
                        torch.empty(B, per, 8, N).pin_memory()]
            s_d2h = torch.cuda.Stream()
            s_main = torch.cuda.current_stream()
            up = bd.ShardedUpload(fm_host.shape, torch.float32, dev)
            fm_free = [None, None]       # the head that read upload slot k has run
            out_free = [None, None]      # the D2H reader of gather slot k is done

            def run(n):
                for i in range(n):
                    k = i % 2
                    fm_d, ev_in = up.upload(fm_host, after_event=fm_free[k])
                    s_main.wait_event(ev_in)
                    if out_free[k] is not None:
                        s_main.wait_event(out_free[k])
                    srcs = self.step(fm_d)
                    fm_free[k] = torch.cuda.Event()
                    fm_free[k].record(s_main)
                    with torch.cuda.stream(s_d2h):
                        s_d2h.wait_event(fm_free[k])
                        if self.last_event is not None:
                            s_d2h.wait_event(self.last_event)      # K3 ran on its side stream
                        for dst, src in zip(out_host, srcs):
                            dst.copy_(src, non_blocking=True)
                            src.record_stream(s_d2h)
                        out_free[k] = torch.cuda.Event()
                        out_free[k].record(s_d2h)
                self.drain()
                s_main.wait_stream(s_d2h)
                s_main.wait_stream(up.slots.stream if up.slots is not None else up.stream)

            run(max(2, warmup))
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            run(steps)
            f1.record()
            barrier()
            ms = max_over_ranks(f0.elapsed_time(f1))
            return {"value": B * self.C_global * steps / (ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": up.bytes_per_rank(), "d2h_bytes_per_step": sum(t.numel() for t in out_host) * 4,
                    "ms_per_step": ms / steps,
                    "note": "bytes are per rank; the replicated feature map is uploaded 1/N per rank and exchanged over NVLink"
                            if world > 1 else "feature map uploaded from pinned host memory every step"}

        def e2e_detections(self, steps, warmup, coder, img):
            """The user-level pipeline end to end with HOST buffers: feature map from pinned host memory (sharded upload at
            N > 1) -> head -> decode + per-label NMS on this rank's labels -> gather of the survivors -> the detections
            (boxes, scores, labels, default boxes, corners: what evaluate.py:118 moves to the CPU) copied to pinned host memory.
            More work per step than `e2e` (it includes the post-processing), far fewer bytes back to the host."""
            det = bd.ClassShardedDetector([None] * self.C_global, list(range(self.C_global)), lambda block: self.head, coder)
            kw = dict(nms_score_threshold=float("-inf"), nms_iou_threshold=0.3)
            up = bd.ShardedUpload(fm_host.shape, torch.float32, dev)
            host_rows = torch.empty(400000, 18).pin_memory()
            s_main = torch.cuda.current_stream()
            fm_free = [None, None]
            stats = {"rows": 0}

            def to_host(dets):
                rows = torch.cat([dets.bbox_xyxy, dets.get_field("scores")[:, None],
                                  dets.get_field("labels").to(torch.int32).view(torch.float32)[:, None],
                                  dets.get_field("default_boxes").bbox_xyxy, dets.get_field("transform_corners")], dim=1)
                t = min(rows.shape[0], host_rows.shape[0])
                host_rows[:t].copy_(rows[:t], non_blocking=True)
                stats["rows"] = t

            def run(n):
                nxt = up.upload(fm_host, after_event=fm_free[0])
                pend = None
                for i in range(n):
                    fm_d, ev_in = nxt
                    if i + 1 < n:                                  # prefetch: the next upload overlaps this step's kernels
                        nxt = up.upload(fm_host, after_event=fm_free[(i + 1) % 2])
                    s_main.wait_event(ev_in)
                    with torch.no_grad():
                        cur = det.submit([fm_d[:1]], [img], **kw)
                    fm_free[i % 2] = torch.cuda.Event()
                    fm_free[i % 2].record(s_main)
                    if pend is not None:                           # finish image i-1 while image i runs
                        to_host(det.result(pend))
                    pend = cur
                to_host(det.result(pend))
                torch.cuda.synchronize()

            run(max(2, warmup))
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            run(steps)
            f1.record()
            barrier()
            ms = max_over_ranks(f0.elapsed_time(f1))
            return {"value": self.C_global * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
                    "h2d_bytes_per_step": up.bytes_per_rank(), "d2h_bytes_per_step": stats["rows"] * 72,
                    "detections": stats["rows"],
                    "what": "host feature map -> head -> decode + per-label NMS (sharded by label) -> survivors gathered -> "
                            "detections in pinned host memory; score threshold -inf; one image per step"}

        def parity(self):
            """N > 1: recompute the class block of rank (r + 1) % N on this rank and compare it bit for bit with the slice
            the gather delivered (driver-visible proof that the multi-GPU result equals the single-GPU one)."""
            if self.sharded is None:
                return None
            with torch.no_grad():
                h = self.sharded.submit(fm_dev)
                loc, score, corners = h.wait()
                other = (rank + 1) % world
                cls = synth_classes(self.per, self.seed0 + other).to(dev)
                ref = hc.create_os2d_head([cls[i:i + 1] for i in range(self.per)])
                rloc, rscore, _, rcorners = ref(fm_dev)
                lo, hi = other * self.per, (other + 1) * self.per
                ok = bool(torch.equal(loc[:, lo:hi], rloc) and torch.equal(score[:, lo:hi], rscore)
                          and torch.equal(corners[:, lo:hi], rcorners))
                torch.cuda.synchronize()
            self.sharded.drain()
            return {"gathered_equals_recomputed": all_true(ok),
                    "checked": "every rank recomputed the {}-class block of rank (r+1)%{} with a fresh single-GPU head and "
                               "torch.equal'ed loc/score/corners against its gathered slice".format(self.per, world)}

    # ---- headline: weak scaling, configs[1] per GPU ----
    weak = Workload(args.classes, CLASS_SEED, args.gather)
    C = args.classes
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total, stage_avg, launches = weak.timed(args.steps, args.warmup, with_stages=True, sampler=sampler)
    clocks = sampler.stop() if rank == 0 else None
    value = B * C * world * args.steps / (ms_total * 1e-3)
    gather_used = weak.sharded.transport() if weak.sharded is not None else None
    e2e = weak.e2e(args.steps, args.warmup)
    parity = weak.parity()
    e2e_det = None
    if B == 1 and not args.no_pipeline:
        try:
            coder0 = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, hc.box_grid_generator_image_level,
                                  lambda s: FeatureMapSize(w=-(-s.w // 16), h=-(-s.h // 16)))
            e2e_det = weak.e2e_detections(args.steps, args.warmup, coder0, FeatureMapSize(w=fm_side * 16, h=fm_side * 16))
        except Exception as e:   # noqa: BLE001
            e2e_det = {"error": repr(e)}

    # ---- host link: device -> host rate of one rank alone and of all ranks at once (the e2e download of the score maps is
    # bound by the second figure on boxes whose GPUs share the host path) ----
    host_link = None
    if world > 1:
        buf_d = torch.empty(32 * 1024 * 1024 // 4, device=dev)
        buf_h = torch.empty(32 * 1024 * 1024 // 4).pin_memory()

        def d2h_rate(active):
            barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            if active:
                for _ in range(10):
                    buf_h.copy_(buf_d, non_blocking=True)
            t1.record()
            barrier()
            ms = max_over_ranks(t0.elapsed_time(t1) if active else 0.0)
            return 10 * buf_d.numel() * 4 / (ms * 1e-3) / 1e9
        alone = d2h_rate(rank == 0)
        together = d2h_rate(True)
        host_link = {"d2h_gbs_one_rank_alone": alone, "d2h_gbs_per_rank_all_ranks_at_once": together,
                     "d2h_gbs_aggregate": together * world}

    # ---- sustained: the same step looped for >= args.sustained_seconds with its own clocks record ----
    sustained = None
    if args.sustained_seconds > 0:
        n_sus = max(args.steps, int(args.sustained_seconds * 1e3 / (ms_total / args.steps)) + 1)
        s2 = ClockSampler(local_rank) if rank == 0 else None
        ms_sus, stage_sus, _ = weak.timed(n_sus, 1, with_stages=True, sampler=s2)
        sustained = {"value": B * C * world * n_sus / (ms_sus * 1e-3), "unit": UNIT, "steps": n_sus, "seconds": ms_sus * 1e-3,
                     "ms_per_step": ms_sus / n_sus, "stage_ms": stage_sus, "clocks": s2.stop() if rank == 0 else None}

    # ---- pipeline: head + decode + per-label NMS on this rank's labels (+ gather of the survivors) ----
    pipeline = None
    if not args.no_pipeline:
        try:
            coder = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, hc.box_grid_generator_image_level,
                                 lambda s: FeatureMapSize(w=-(-s.w // 16), h=-(-s.h // 16)))
            own = synth_classes(C, CLASS_SEED + rank).to(dev)
            maps = [None] * (C * world)
            for i in range(C):
                maps[rank * C + i] = own[i:i + 1]
            img = FeatureMapSize(w=fm_side * 16, h=fm_side * 16)
            with torch.no_grad():
                det = bd.ClassShardedDetector(maps, list(range(C * world)), hc.create_os2d_head, coder)
                kw = dict(nms_score_threshold=float("-inf"), nms_iou_threshold=0.3)     # reference defaults, config.py:198-200
                for _ in range(max(2, args.warmup)):
                    dets = det([fm_dev[:1]], [img], **kw)
                barrier()
                l0 = lib.os2d_b200_launch_count()
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
                pend = None
                for _ in range(args.steps):                    # image i+1 is submitted before the result of image i is read
                    nxt = det.submit([fm_dev[:1]], [img], **kw)
                    if pend is not None:
                        dets = det.result(pend)
                    pend = nxt
                dets = det.result(pend)
                p1.record()
                barrier()
                pp_ms = max_over_ranks(p0.elapsed_time(p1)) / args.steps
                # decode + NMS alone on this rank's outputs (launch count and time of the post-processing itself)
                loc, score, _, corners = det.head(fm_dev[:1])
                a = ([loc[0].view(C, 4, N)], [score[0].view(C, N)], [img], det.class_ids)
                kw2 = dict(kw, transform_corners_pyramid=[corners[0].view(C, 8, N)])
                coder.decode_pyramid(*a, **kw2)
                torch.cuda.synchronize()
                l1 = lib.os2d_b200_launch_count()
                q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                q0.record()
                for _ in range(5):
                    local = coder.decode_pyramid(*a, **kw2)
                q1.record()
                torch.cuda.synchronize()
                pp_launches = (lib.os2d_b200_launch_count() - l1) // 5
            pipeline = {"value": C * world / (pp_ms * 1e-3), "unit": UNIT, "ms_per_step": pp_ms,
                        "decode_nms_ms_per_image": q0.elapsed_time(q1) / 5, "decode_nms_launches": pp_launches,
                        "candidates_per_gpu": C * N, "detections": len(dets), "detections_this_rank": len(local),
                        "what": "head + decode + per-label NMS sharded by label + gather of the survivors, score threshold -inf "
                                "(reference default: every anchor is an NMS candidate), one image per step, timed on the device"}
        except Exception as e:   # noqa: BLE001
            pipeline = {"error": repr(e)}

    # ---- strong scaling: configs[2], 1000 classes sharded ceil(1000/N) per GPU ----
    strong = None
    if args.strong_classes > 0:
        try:
            del weak
            torch.cuda.empty_cache()
            per = -(-args.strong_classes // world)
            sw = Workload(per, 4321, args.gather)
            ms_s, stage_s, _ = sw.timed(args.steps, args.warmup, with_stages=True)
            strong = {"value": B * per * world * args.steps / (ms_s * 1e-3), "unit": UNIT, "scaling": "strong",
                      "global_classes": per * world, "classes_per_gpu": per, "ms_per_step": ms_s / args.steps, "stage_ms": stage_s,
                      "e2e": sw.e2e(args.steps, args.warmup), "parity": sw.parity(),
                      "workload": "configs[2]: 1280px synthetic input, 1000 query classes sharded over {} GPU(s), gather of the "
                                  "per-class outputs before NMS".format(world)}
            del sw
        except Exception as e:   # noqa: BLE001
            strong = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (conv1 of the TransformNet) + the correlation GEMM ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:   # noqa: BLE001
        pass
    # the 20-step timed region lasts ~40 ms: the clocks stay at the burst level, so the denominator is the BURST peak;
    # the `sustained` block is measured over seconds and is compared with the sustained peak
    peak_burst = peaks.get("bf16_tflops") or 1640.0
    peak_sus = peaks.get("bf16_tflops_sustained") or 1400.0
    src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    planes = B * C
    flops = {"conv1": 2.0 * N * 225 * 128 * 49 * planes, "corr": 2.0 * N * 225 * D * planes,
             "conv2": 2.0 * N * 128 * 64 * 25 * planes, "conv3": 2.0 * N * 64 * 6 * 25 * planes}
    traffic = {}
    try:
        if B == 1 and C == 100 and fm_side == 80:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:   # noqa: BLE001
        traffic = {}
    kernel_names = {"conv1": "conv::conv_kernel<7>", "conv2": "conv::conv_kernel<5>", "corr": "corr::corr_kernel<1>",
                    "conv3": "conv3s::conv3s_kernel"}

    def roof(name, stages, peak, peak_name):
        ms = stages.get(name)
        if not ms:
            return None
        ach = flops[name] / (ms * 1e-3) / 1e12
        tr = traffic.get(kernel_names.get(name, ""), {}).get("traffic_bytes")
        return {"kernel": kernel_names.get(name, name), "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": tr, "traffic_unit": "bytes per launch (ncu dram read+write)",
                "algorithmic_flop_per_launch": flops[name], "ms_per_launch": ms,
                "peak_source": "{} {} (fp16 and bf16 share the tensor rate)".format(src, peak_name)}

    roofline = roof("conv1", stage_avg, peak_burst, "bf16_tflops (burst: the timed region is tens of ms)")
    total_flop = sum(flops.values())
    extra = {"roofline_corr": roof("corr", stage_avg, peak_burst, "bf16_tflops (burst)"),
             "roofline_conv2": roof("conv2", stage_avg, peak_burst, "bf16_tflops (burst)"),
             "roofline_conv3": roof("conv3", stage_avg, peak_burst, "bf16_tflops (burst)"),
             "stage_ms": stage_avg,
             "tensor_frac_whole_step": total_flop * args.steps / (ms_total * 1e-3) / 1e12 / peak_burst}
    if sustained is not None:
        sustained["roofline"] = roof("conv1", sustained["stage_ms"], peak_sus, "bf16_tflops_sustained")
        sustained["tensor_frac_whole_step"] = total_flop * sustained["steps"] / (sustained["seconds"]) / 1e12 / peak_sus

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        sc = max(1, min(args.cpu_sample_classes, C))
        rate, _, sample, used, kind = cpu_reference_rate(args, 3, 1, sc)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": used, "kind": kind, "sample": sample}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 operands, fp32 accumulate (hi/lo-split weights)", "data": "synthetic",
            "config": dict(workload_config(args, world), gather=(gather_used if world > 1 else None),
                           wave_classes=(args.wave_classes or None),
                           concurrent_corr_sms=(args.concurrent_corr if args.concurrent_corr >= 0 else "library default"),
                           async_resample=((args.async_resample == 1) if args.async_resample >= 0 else ASYNC_RESAMPLE_DEFAULT)),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "parity": parity, "strong_c1000": strong, "pipeline": pipeline, "sustained": sustained,
            "e2e_detections": e2e_det, "host_link": host_link}
    line.update(extra)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
