"""Stage-by-stage numerical check of the CUDA kernels against torch ops on the same GPU (diagnostics tool).

    python tests/tools/gpu_stage_check.py <stage> [H W C B]
stages: env pack corr conv1 conv2 conv3 resample head decode all
Each stage prints max abs / rel errors; exit code 1 if a stage is out of tolerance.
Run every stage under `timeout` (a wrong barrier protocol hangs rather than fails).
"""
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # repository root
import torch
import torch.nn.functional as F

from os2d_b200 import _cabi
from os2d_b200 import head as bh
from os2d_b200.structures import FeatureMapSize
from oracle import head_oracle as ho

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

DEV = "cuda"


def relerr(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item(), (a - b).abs().max().item()


def report(name, a, b, tol):
    r, ab = relerr(a, b)
    ok = r <= tol and bool(torch.isfinite(a.float()).all())
    print("  [{}] {:<28s} rel-to-max {:.3e} abs {:.3e} (tol {:.1e})".format("OK" if ok else "FAIL", name, r, ab, tol), flush=True)
    return ok


def make_inputs(H, W, C, B, seed=0, D=1024):
    g = torch.Generator().manual_seed(seed)
    sizes = [(15, 15), (12, 18), (19, 11)]
    cms = [(torch.randn(1, D, *sizes[i % 3], generator=g) * 0.5 + 0.2).relu() for i in range(C)]
    fm = (torch.randn(B, D, H, W, generator=g) * 0.5 + 0.2).relu()
    return cms, fm


def unpack_z(zvol, planes, H, W):
    """[planes,30,N,8] fp16 -> [planes,240,H,W] fp32 (raw stored values)."""
    return zvol.float().view(planes, 30, H * W, 8).permute(0, 1, 3, 2).reshape(planes, 240, H, W)


def unpack_vol(vol, planes, chunks, H, W):
    return vol.float().view(planes, chunks, H * W, 8).permute(0, 1, 3, 2).reshape(planes, chunks * 8, H, W)


def pack_vol(x):
    """[planes,Cc,H,W] fp32 -> [planes,Cc/8,N,8] fp16"""
    p, c, H, W = x.shape
    return x.view(p, c // 8, 8, H * W).permute(0, 1, 3, 2).contiguous().to(torch.float16)


def stage_env():
    print(torch.__version__, torch.cuda.get_device_name(0))
    pr = torch.cuda.get_device_properties(0)
    print("SMs", pr.multi_processor_count, "L2", getattr(pr, "L2_cache_size", None), "mem", pr.total_memory)
    lib = _cabi.load()
    print("abi", lib.os2d_b200_abi_version(), "num_sms", lib.os2d_b200_num_sms())
    print("cpu count", os.cpu_count())
    return True


def run_k0k1(H, W, C, B):
    lib = _cabi.load()
    cms, fm = make_inputs(H, W, C, B)
    cms_d = [c.to(DEV) for c in cms]
    fm_d = fm.to(DEV)
    cf32, packed = bh._prepare_class_operands(cms_d)
    N = H * W
    D = fm.shape[1]
    img_packed = torch.empty(B, N, D, dtype=torch.float16, device=DEV)
    inv_ws = torch.empty(B, N, dtype=torch.float32, device=DEV)
    st = _cabi.stream_ptr()
    _cabi.check(lib.os2d_pack_image_features(_cabi.ptr(fm_d), B, D, N, _cabi.ptr(inv_ws), _cabi.ptr(img_packed), st), "pack_image")
    return cms, fm, cf32, packed, img_packed


def stage_pack(H, W, C, B):
    cms, fm, cf32, packed, img_packed = run_k0k1(H, W, C, B)
    torch.cuda.synchronize()
    ok = True
    cf_ref = ho.prepare_class_features(cms)
    ok &= report("class cf32 vs oracle", cf32.cpu(), cf_ref, 1e-5)
    pk_ref = cf_ref.permute(0, 3, 2, 1).reshape(C, 225, -1) * 32
    ok &= report("class packed[:225]", packed[:, :225].float().cpu(), pk_ref, 1e-3)
    ok &= report("class packed pad rows == 0", packed[:, 225:].float().abs().max().cpu() + 1, torch.ones(()), 1e-9)
    f = ho.l2_normalize(fm, 1e-5)
    ref = f.reshape(B, f.shape[1], -1).permute(0, 2, 1) * 32
    ok &= report("image packed", img_packed.float().cpu(), ref, 1e-3)
    return ok


def stage_corr(H, W, C, B):
    lib = _cabi.load()
    cms, fm, cf32, packed, img_packed = run_k0k1(H, W, C, B)
    N = H * W
    planes = B * C
    zvol = torch.zeros(planes, 30, N, 8, dtype=torch.float16, device=DEV)
    rawvol = torch.zeros(planes, 225, N, dtype=torch.float16, device=DEV)
    t0 = time.time()
    _cabi.check(lib.os2d_correlate(_cabi.ptr(img_packed), _cabi.ptr(packed), B, C, fm.shape[1], H, W, _cabi.ptr(zvol),
                                   _cabi.ptr(rawvol), _cabi.stream_ptr()), "correlate")
    torch.cuda.synchronize()
    print("  corr kernel returned in {:.3f}s".format(time.time() - t0))
    # reference from the very same fp16 operands
    a = packed[:, :225].float()                     # [C,225,D]
    b = img_packed.float()                          # [B,N,D]
    corr = torch.einsum("ckd,bnd->bckn", a, b) / 1024.0    # [B,C,225,N]
    corr = corr.reshape(planes, 225, N)
    ok = report("raw corr", rawvol.float(), corr, 1e-3)
    z = F.relu(corr)
    z = z / (z.pow(2).sum(1, keepdim=True).sqrt() + 1e-6)
    m = z.mean(1, keepdim=True)
    zc = (z - m) * 64
    zu = zvol.float().permute(0, 1, 3, 2).reshape(planes, 240, N)
    ok &= report("z centred (ch<225)", zu[:, :225], zc, 2e-3)
    m8 = (m * 8).squeeze(1)
    ok &= report("dc hi+lo == 8*mean", zu[:, 225] + zu[:, 227], m8, 1e-5)
    ok &= report("dc ch226 == ch225", zu[:, 226], zu[:, 225], 1e-9)
    ok &= report("pad channels zero", zu[:, 228:].abs().max() + 1, torch.ones((), device=DEV), 1e-9)
    return ok


def _tn(P, seed=1, spread=0.005):
    return ho.random_transform_net(P, seed=seed, spread=spread)


def stage_conv(layer, H, W, C, B):
    lib = _cabi.load()
    P = 6
    tn = _tn(P)
    pw = bh.pack_transform_net(tn, P, DEV)
    planes = B * C
    N = H * W
    g = torch.Generator().manual_seed(3)
    st = _cabi.stream_ptr()
    if layer == 1:
        # synthetic z: positive, roughly unit-norm over 225 channels
        z = torch.rand(planes, 225, H, W, generator=g) + 0.5
        z = z / z.pow(2).sum(1, keepdim=True).sqrt()
        m = z.mean(1, keepdim=True)
        zc = ((z - m) * 64).to(torch.float16)
        m8 = m * 8
        mh = m8.to(torch.float16)
        ml = (m8 - mh.float()).to(torch.float16)
        full = torch.zeros(planes, 240, H, W, dtype=torch.float16)
        full[:, :225] = zc
        full[:, 225:226] = mh
        full[:, 226:227] = mh
        full[:, 227:228] = ml
        vol = pack_vol(full.float()).to(DEV)
        out = torch.zeros(planes, 16, N, 8, dtype=torch.float16, device=DEV)
        t0 = time.time()
        _cabi.check(lib.os2d_transform_conv(1, 128, _cabi.ptr(vol), _cabi.ptr(pw["w1"]), _cabi.ptr(pw["alpha1"]),
                                            _cabi.ptr(pw["beta1"]), _cabi.ptr(out), planes, H, W, st), "conv1")
        torch.cuda.synchronize()
        print("  conv1 kernel returned in {:.3f}s".format(time.time() - t0))
        zq = (zc.float() / 64 + (mh.float() + ml.float()) / 8).to(DEV)
        a1, b1 = ho.fold_bn(tn["conv.0.weight"], tn["conv.0.bias"], tn["conv.1.weight"], tn["conv.1.bias"],
                            tn["conv.1.running_mean"], tn["conv.1.running_var"])
        ref = F.conv2d(zq.double(), tn["conv.0.weight"].double().to(DEV), None, padding=3)
        ref = F.relu(ref * a1.double().to(DEV).view(1, -1, 1, 1) + b1.double().to(DEV).view(1, -1, 1, 1)).float()
        got = unpack_vol(out, planes, 16, H, W)
        return report("conv1 out", got, ref, 2e-3)
    if layer == 2:
        h1 = torch.rand(planes, 128, H, W, generator=g).to(torch.float16)
        vol = pack_vol(h1.float()).to(DEV)
        out = torch.zeros(planes, 16, N, 8, dtype=torch.float16, device=DEV)
        t0 = time.time()
        _cabi.check(lib.os2d_transform_conv(2, 64, _cabi.ptr(vol), _cabi.ptr(pw["w2"]), _cabi.ptr(pw["alpha2"]),
                                            _cabi.ptr(pw["beta2"]), _cabi.ptr(out), planes, H, W, st), "conv2")
        torch.cuda.synchronize()
        print("  conv2 kernel returned in {:.3f}s".format(time.time() - t0))
        a2, b2 = ho.fold_bn(tn["conv.3.weight"], tn["conv.3.bias"], tn["conv.4.weight"], tn["conv.4.bias"],
                            tn["conv.4.running_mean"], tn["conv.4.running_var"])
        ref = F.conv2d(h1.double().to(DEV), tn["conv.3.weight"].double().to(DEV), None, padding=2)
        ref = F.relu(ref * a2.double().to(DEV).view(1, -1, 1, 1) + b2.double().to(DEV).view(1, -1, 1, 1)).float()
        full = unpack_vol(out, planes, 16, H, W)
        got = full[:, :64] + full[:, 64:]             # fp16 value planes + fp16 residual planes
        ok = report("conv2 out (hi only)", full[:, :64], ref, 2e-3)
        return report("conv2 out (hi+lo)", got, ref, 1e-5) and ok
    if layer == 3:
        h2f = torch.rand(planes, 64, H, W, generator=g) * 3
        h2hi = h2f.to(torch.float16)
        h2lo = (h2f - h2hi.float()).to(torch.float16)
        h2 = h2hi.double() + h2lo.double()
        vol = pack_vol(torch.cat([h2hi.float(), h2lo.float()], dim=1)).to(DEV)
        out = torch.zeros(planes, P, N, dtype=torch.float32, device=DEV)
        t0 = time.time()
        _cabi.check(lib.os2d_transform_conv(3, P, _cabi.ptr(vol), _cabi.ptr(pw["w3"]), _cabi.ptr(pw["alpha3"]),
                                            _cabi.ptr(pw["beta3"]), _cabi.ptr(out), planes, H, W, st), "conv3")
        torch.cuda.synchronize()
        print("  conv3 kernel returned in {:.3f}s".format(time.time() - t0))
        ref = F.conv2d(h2.double().to(DEV), tn["linear.weight"].double().to(DEV), tn["linear.bias"].double().to(DEV),
                       padding=2).float()
        return report("conv3 out", out.view(planes, P, H, W), ref, 1e-4)


def stage_resample(H, W, C, B):
    lib = _cabi.load()
    planes = B * C
    N = H * W
    g = torch.Generator().manual_seed(5)
    ok = True
    for P, inverse in ((6, 1), (6, 0), (4, 1), (4, 0)):
        corr = torch.rand(planes, 225, H, W, generator=g).to(torch.float16).float()
        ident = torch.tensor([1., 0, 0, 0, 1, 0]) if P == 6 else torch.tensor([1., 0, 1, 0])
        params = ident.view(1, P, 1, 1) + 0.15 * torch.randn(planes, P, H, W, generator=g)
        theta = ho.theta_from_params(params, P == 4, bool(inverse))
        score_ref = ho.resample_and_pool(corr, theta)
        loc_ref, cor_ref = ho.boxes_and_corners(theta, H, W)
        raw = corr.reshape(planes, 225, N).to(torch.float16).to(DEV)
        pr = params.reshape(planes, P, N).contiguous().to(DEV)
        score = torch.zeros(planes, N, device=DEV)
        loc = torch.zeros(planes, 4, N, device=DEV)
        cor = torch.zeros(planes, 8, N, device=DEV)
        _cabi.check(lib.os2d_resample_boxes(_cabi.ptr(raw), _cabi.ptr(pr), planes, P, H, W, inverse, 16.0, 16.0, 240.0, 240.0,
                                            _cabi.ptr(score), _cabi.ptr(loc), _cabi.ptr(cor), N, 4 * N, 8 * N,
                                            _cabi.stream_ptr()), "resample")
        torch.cuda.synchronize()
        tag = "P{} inv{}".format(P, inverse)
        ok &= report("score " + tag, score.cpu().view(planes, H, W), score_ref, 1e-4)
        ok &= report("loc " + tag, loc.cpu().view(planes, 4, H, W), loc_ref, 1e-4)
        ok &= report("corners " + tag, cor.cpu().view(planes, 8, H, W), cor_ref, 1e-5)
    return ok


def stage_head(H, W, C, B):
    ok = True
    for simple, inverse in ((False, True), (True, False), (False, False), (True, True)):
        P = 4 if simple else 6
        tn = _tn(P)
        cms, fm = make_inputs(H, W, C, B)
        hc = bh.build_os2d_head_creator(simple, True, inverse, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
        sd = {k: v for k, v in tn.items()}
        hc.aligner.parameter_regressor.load_state_dict(sd, strict=False)
        hc.eval()
        with torch.no_grad():
            head = hc.create_os2d_head([c.to(DEV) for c in cms])
            t0 = time.time()
            loc, rec, rec2, corners = head(fm.to(DEV))
            torch.cuda.synchronize()
            dt = time.time() - t0
            cf = ho.prepare_class_features(cms)
            oloc, osc, ocor = ho.head_forward(cf, fm, tn, simple, inverse)
        tag = "simple{} inv{}".format(int(simple), int(inverse))
        print("  head forward {:.3f}s".format(dt))
        ok &= report("score " + tag, rec.cpu(), osc, 1e-3)
        ok &= report("loc " + tag, loc.cpu(), oloc, 1e-3)
        ok &= report("corners " + tag, corners.cpu(), ocor, 1e-3)
    return ok


def main():
    stage = sys.argv[1] if len(sys.argv) > 1 else "all"
    H, W, C, B = [int(v) for v in (sys.argv[2:6] if len(sys.argv) >= 6 else (20, 27, 3, 2))]
    print("== stage {} H{} W{} C{} B{}".format(stage, H, W, C, B), flush=True)
    fn = {
        "env": lambda: stage_env(),
        "pack": lambda: stage_pack(H, W, C, B),
        "corr": lambda: stage_corr(H, W, C, B),
        "conv1": lambda: stage_conv(1, H, W, C, B),
        "conv2": lambda: stage_conv(2, H, W, C, B),
        "conv3": lambda: stage_conv(3, H, W, C, B),
        "resample": lambda: stage_resample(H, W, C, B),
        "head": lambda: stage_head(H, W, C, B),
    }[stage]
    ok = fn()
    print("== stage {} {}".format(stage, "PASSED" if ok else "FAILED"), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
