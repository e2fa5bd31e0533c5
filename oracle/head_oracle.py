"""CPU restatement (torch fp32, explicit arithmetic) of the OS2D head hot path.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.  Every function cites the reference
lines (relative to /root/reference) whose arithmetic it restates.  The restatement is
independent code: explicit bilinear gathers instead of F.grid_sample, closed-form affine
inverse instead of torch.inverse, analytic anchor grids instead of cached box tensors.

An optional ``emulate`` mode rounds the same intermediates the CUDA path keeps in fp16
(MMA operands, stored correlation volume, hidden activations) so the tolerance budget of the
GPU kernels can be studied on a CPU-only machine.
"""
import math

import torch
import torch.nn.functional as F

GRID = 15            # os2d/modeling/head.py:66-69 (out grid / reference fm size / receptive field)
NUM_CH = GRID * GRID  # 225 correlation channels
POOL_BORDER = 2      # os2d/modeling/head.py:280
BOX_WEIGHTS = (10.0, 10.0, 5.0, 5.0)  # os2d/modeling/box_coder.py:13
BN_EPS = 1e-5        # nn.BatchNorm2d default used by head.py:623

# power-of-two operand pre-scales used by the CUDA path (emulation only)
SCALE_FEAT = 32.0
SCALE_Z = 16.0


def _q16(x):
    """Round to fp16 and back (emulation of fp16 storage / MMA operands)."""
    return x.to(torch.float16).to(torch.float32)


def l2_normalize(x, eps):
    """x / (||x||_2 over dim 1 + eps).  os2d/modeling/head.py:597-601 (eps added to the norm)."""
    return x / (x.pow(2).sum(dim=1, keepdim=True).sqrt() + eps)


def resize_class_map(fm):
    """Bilinear resize of one class map [1,D,h,w] to [1,D,15,15], align_corners=True, zero pad.
    os2d/modeling/head.py:241-259 (identity affine_grid + grid_sample)."""
    _, d, h, w = fm.shape
    # F.affine_grid builds the normalised grid -1 + 2 i/14 in fp32 and grid_sample un-normalises
    # ((g + 1) / 2) * (size - 1); restate that rounding order.
    lin = torch.linspace(-1, 1, GRID, dtype=torch.float32)
    ys = (lin + 1) / 2 * (h - 1)
    xs = (lin + 1) / 2 * (w - 1)
    y0 = ys.floor()
    x0 = xs.floor()
    wy1 = ys - y0
    wx1 = xs - x0
    wy0 = 1 - wy1
    wx0 = 1 - wx1
    y0 = y0.long()
    x0 = x0.long()
    y1 = y0 + 1
    x1 = x0 + 1

    def tap(yi, xi):
        valid_y = (yi >= 0) & (yi <= h - 1)
        valid_x = (xi >= 0) & (xi <= w - 1)
        v = fm[0][:, yi.clamp(0, h - 1)][:, :, xi.clamp(0, w - 1)]  # [D,15,15]
        m = (valid_y[:, None] & valid_x[None, :]).to(fm.dtype)
        return v * m

    out = (tap(y0, x0) * (wy0[:, None] * wx0[None, :]) + tap(y0, x1) * (wy0[:, None] * wx1[None, :])
           + tap(y1, x0) * (wy1[:, None] * wx0[None, :]) + tap(y1, x1) * (wy1[:, None] * wx1[None, :]))
    return out.unsqueeze(0)


def prepare_class_features(class_maps):
    """list of [1,D,h_c,w_c] -> L2-normalised [C,D,15,15].  head.py:261-268, 293."""
    cf = torch.cat([resize_class_map(m) for m in class_maps], dim=0)
    return l2_normalize(cf, 1e-5)


def correlate(cf, f, emulate=False):
    """corr[b*C+c, tx*15+ty, y, x] = sum_d cf[c,d,ty,tx] f[b,d,y,x].  head.py:342-350
    (note the transposed channel order k = tx*15 + ty)."""
    C, D = cf.shape[:2]
    B, _, H, W = f.shape
    a = cf.permute(0, 3, 2, 1).reshape(C * NUM_CH, D)   # row (c, tx, ty)
    b = f.reshape(B, D, H * W)
    if emulate:
        a = _q16(a * SCALE_FEAT) / SCALE_FEAT
        b = _q16(b * SCALE_FEAT) / SCALE_FEAT
    corr = torch.matmul(a.unsqueeze(0), b)             # [B, C*225, N]
    return corr.reshape(B * C, NUM_CH, H, W)


def fold_bn(conv_w, conv_b, bn_w, bn_b, bn_mean, bn_var):
    """Eval-mode BatchNorm folded into per-output-channel scale/shift.  head.py:619-627."""
    alpha = bn_w / torch.sqrt(bn_var + BN_EPS)
    beta = (conv_b - bn_mean) * alpha + bn_b
    return alpha, beta


def transform_net(corr, tn, emulate=False):
    """ReLU -> L2norm(225, eps 1e-6) -> conv7/BN/ReLU -> conv5/BN/ReLU -> conv5.  head.py:648-655.
    ``tn`` is a state dict with the reference key names (conv.0/1/3/4, linear)."""
    z = l2_normalize(F.relu(corr), 1e-6)
    if not emulate:
        h = F.conv2d(z, tn["conv.0.weight"], tn["conv.0.bias"], padding=3)
        h = F.batch_norm(h, tn["conv.1.running_mean"], tn["conv.1.running_var"],
                         tn["conv.1.weight"], tn["conv.1.bias"], False, 0.0, BN_EPS)
        h = F.relu(h)
        h = F.conv2d(h, tn["conv.3.weight"], tn["conv.3.bias"], padding=2)
        h = F.batch_norm(h, tn["conv.4.running_mean"], tn["conv.4.running_var"],
                         tn["conv.4.weight"], tn["conv.4.bias"], False, 0.0, BN_EPS)
        h = F.relu(h)
        return F.conv2d(h, tn["linear.weight"], tn["linear.bias"], padding=2), z

    # fp16-operand emulation of the CUDA plan: fp16 z, fp16 weights, fp32 accumulate,
    # BN applied as an fp32 epilogue scale/shift, hidden activations stored in fp16.
    zq = _q16(z * SCALE_Z) / SCALE_Z
    a1, b1 = fold_bn(tn["conv.0.weight"], tn["conv.0.bias"], tn["conv.1.weight"], tn["conv.1.bias"],
                     tn["conv.1.running_mean"], tn["conv.1.running_var"])
    h = F.conv2d(zq, _q16(tn["conv.0.weight"]), None, padding=3)
    h = _q16(F.relu(h * a1.view(1, -1, 1, 1) + b1.view(1, -1, 1, 1)))
    a2, b2 = fold_bn(tn["conv.3.weight"], tn["conv.3.bias"], tn["conv.4.weight"], tn["conv.4.bias"],
                     tn["conv.4.running_mean"], tn["conv.4.running_var"])
    h = F.conv2d(h, _q16(tn["conv.3.weight"]), None, padding=2)
    h = _q16(F.relu(h * a2.view(1, -1, 1, 1) + b2.view(1, -1, 1, 1)))
    p = F.conv2d(h, _q16(tn["linear.weight"]), None, padding=2) + tn["linear.bias"].view(1, -1, 1, 1)
    return p, zq


INVERSE_CHUNK = 256 * 256 - 1      # head.py:141 (batched_inverse splits the matrices in chunks of 65535)
INVERSE_REG = 1e-5                  # head.py:129 (added to the diagonal of every matrix of a chunk whose inversion raised)


def theta_from_params(p, simple_affine, inverse):
    """[NB,P,H,W] regressed parameters -> six affine coefficients (a,b,tx,c,d,ty) each [NB,H,W].
    head.py:81-153: P=6 -> [[p0,p1,p2],[p3,p4,p5]], P=4 -> [[p0,0,p1],[0,p2,p3]];
    optional inverse of the 3x3 homogeneous matrix [[a,b,tx],[c,d,ty],[0,0,1]] (closed form here, LU in the reference).
    Failure handling of the reference (head.py:123-146, `robust_inverse` inside `batched_inverse`): the matrices, in
    (n, y, x) order, are inverted in chunks of 65535 (one chunk when there are fewer); when torch.inverse raises for a chunk
    (some matrix has an exactly zero LU pivot = is exactly singular), 1e-5 is added to the diagonal of EVERY matrix of
    that chunk, the homogeneous 1 included, before inverting again.  Restated with "exactly singular" = fp32 a*d - b*c == 0:
      inverse of [[a+e,b,tx],[c,d+e,ty],[0,0,1+e]] = [[A_e^-1, -A_e^-1 t / (1+e)]], evaluated in fp64 for those chunks
    (pinned by tests/golden/theta_singular.npz, generated by the reference on CPU)."""
    if simple_affine:
        assert p.shape[1] == 4
        z = torch.zeros_like(p[:, 0])
        a, b, tx, c, d, ty = p[:, 0], z, p[:, 1], z, p[:, 2], p[:, 3]
    else:
        assert p.shape[1] == 6
        a, b, tx, c, d, ty = (p[:, i] for i in range(6))
    if inverse:
        det = a * d - b * c
        ia, ib, ic, id_ = d / det, -b / det, -c / det, a / det
        itx = -(ia * tx + ib * ty)
        ity = -(ic * tx + id_ * ty)
        singular = (det == 0).reshape(-1)
        if bool(singular.any()):
            n = singular.numel()
            flagged = torch.zeros(n, dtype=torch.bool)
            if n >= INVERSE_CHUNK:
                for s0 in range(0, n, INVERSE_CHUNK):
                    if bool(singular[s0:s0 + INVERSE_CHUNK].any()):
                        flagged[s0:s0 + INVERSE_CHUNK] = True
            else:
                flagged[:] = True
            flagged = flagged.view(det.shape)
            e = INVERSE_REG
            ar, dr = a.double() + e, d.double() + e
            br, cr, txr, tyr = b.double(), c.double(), tx.double(), ty.double()
            detr = ar * dr - br * cr
            ra, rb, rc, rd = dr / detr, -br / detr, -cr / detr, ar / detr
            rtx = -(ra * txr + rb * tyr) / (1.0 + e)
            rty = -(rc * txr + rd * tyr) / (1.0 + e)
            ia = torch.where(flagged, ra.float(), ia)
            ib = torch.where(flagged, rb.float(), ib)
            ic = torch.where(flagged, rc.float(), ic)
            id_ = torch.where(flagged, rd.float(), id_)
            itx = torch.where(flagged, rtx.float(), itx)
            ity = torch.where(flagged, rty.float(), ity)
        a, b, tx, c, d, ty = ia, ib, itx, ic, id_, ity
    return a, b, tx, c, d, ty


def _grid_axis():
    """F.affine_grid base coordinates for size 15, align_corners=True: -1 + 2 i / 14.  head.py:184."""
    return torch.linspace(-1, 1, GRID, dtype=torch.float32)


def resample_and_pool(corr, theta, emulate=False):
    """Score = mean over the inner 11x11 grid points of bilinear samples of channel tx*15+ty at
    (clamp(7.5 gx + x + 0.5, 0, W-1), clamp(7.5 gy + y + 0.5, 0, H-1)).
    head.py:371-395 (local->fm coords, unit normalise, clamp), :439-520 (resample, masked mean)."""
    NB, K, H, W = corr.shape
    a, b, tx, c, d, ty = theta
    lin = _grid_axis()
    gi = torch.arange(POOL_BORDER, GRID - POOL_BORDER)   # only the inner points carry mask weight
    xj = lin[gi]            # x_j, index j = template x
    yi = lin[gi]            # y_i, index i = template y
    # [NB,H,W,i,j]
    gx = a[..., None, None] * xj[None, None, None, None, :] + b[..., None, None] * yi[None, None, None, :, None] + tx[..., None, None]
    gy = c[..., None, None] * xj[None, None, None, None, :] + d[..., None, None] * yi[None, None, None, :, None] + ty[..., None, None]
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, W, 1, 1)
    ys = torch.arange(H, dtype=torch.float32).view(1, H, 1, 1, 1)
    # head.py:36-37 with the stride-1, size-15 box grid (box_coder.py:42-60): x_fm = 7.5 gx + (x + 0.5)
    px = gx * 7.5 + (xs + 0.5)
    py = gy * 7.5 + (ys + 0.5)
    # head.py:381-384: normalise to [-1,1], clamp; grid_sample un-normalises again
    ux = (px / (W - 1) * 2 - 1).clamp(-1, 1)
    uy = (py / (H - 1) * 2 - 1).clamp(-1, 1)
    px = ((ux.double() + 1) / 2 * (W - 1))
    py = ((uy.double() + 1) / 2 * (H - 1))
    x0 = px.floor().clamp(0, W - 1)
    y0 = py.floor().clamp(0, H - 1)
    wx1 = px - x0
    wy1 = py - y0
    x0 = x0.long()
    y0 = y0.long()
    x1 = (x0 + 1).clamp(max=W - 1)
    y1 = (y0 + 1).clamp(max=H - 1)
    # channel of grid point (i=ty, j=tx): k = j*15 + i   (head.py:480)
    k = (gi[None, :] * GRID + gi[:, None]).view(1, 1, 1, gi.numel(), gi.numel()).expand_as(x0)
    src = _q16(corr) if emulate else corr
    flat = src.reshape(NB, K * H * W).double()
    base = k * (H * W)

    def gather(yy, xx):
        idx = (base + yy * W + xx).reshape(NB, -1)
        return torch.gather(flat, 1, idx).view_as(px)

    v = (gather(y0, x0) * (1 - wy1) * (1 - wx1) + gather(y0, x1) * (1 - wy1) * wx1
         + gather(y1, x0) * wy1 * (1 - wx1) + gather(y1, x1) * wy1 * wx1)
    n_inner = (GRID - 2 * POOL_BORDER) ** 2
    # head.py:509-519: cast to float, multiply by the 1/121 mask, sum
    score = (v.float() * (1.0 / n_inner)).sum(dim=(-1, -2))
    return score  # [NB,H,W]


def boxes_and_corners(theta, H, W, stride=16, box=240):
    """Box = bbox of the 225 transformed grid points in image coordinates, encoded against the
    anchor; corners = grid points (0,0),(0,14),(14,0),(14,14).  head.py:404-433,
    box_coder.py:306-317 (clip_to_min_size 1, encode_boxes weights 10,10,5,5)."""
    a, b, tx, c, d, ty = theta
    lin = _grid_axis()
    gx = a[..., None, None] * lin[None, None, None, None, :] + b[..., None, None] * lin[None, None, None, :, None] + tx[..., None, None]
    gy = c[..., None, None] * lin[None, None, None, None, :] + d[..., None, None] * lin[None, None, None, :, None] + ty[..., None, None]
    acx = (torch.arange(W, dtype=torch.float32) + 0.5) * stride
    acy = (torch.arange(H, dtype=torch.float32) + 0.5) * stride
    half = box / 2.0
    X = gx * half + acx.view(1, 1, W, 1, 1)
    Y = gy * half + acy.view(1, H, 1, 1, 1)
    NB = X.shape[0]
    Xf = X.reshape(NB, H, W, -1)
    Yf = Y.reshape(NB, H, W, -1)
    x1 = Xf.min(-1)[0]
    y1 = Yf.min(-1)[0]
    x2 = Xf.max(-1)[0]
    y2 = Yf.max(-1)[0]
    x2 = torch.where(x1 + 1 > x2, x1 + 1, x2)     # bounding_box.py:267-277
    y2 = torch.where(y1 + 1 > y2, y1 + 1, y2)
    # torchvision encode_boxes(reference=class box, proposals=anchor)
    ax1 = acx.view(1, 1, W) - half
    ay1 = acy.view(1, H, 1) - half
    aw = torch.full((1, 1, 1), float(box))
    ah = torch.full((1, 1, 1), float(box))
    actr_x = ax1 + 0.5 * aw
    actr_y = ay1 + 0.5 * ah
    gw = x2 - x1
    gh = y2 - y1
    gcx = x1 + 0.5 * gw
    gcy = y1 + 0.5 * gh
    loc = torch.stack([BOX_WEIGHTS[0] * (gcx - actr_x) / aw,
                       BOX_WEIGHTS[1] * (gcy - actr_y) / ah,
                       BOX_WEIGHTS[2] * torch.log(gw / aw),
                       BOX_WEIGHTS[3] * torch.log(gh / ah)], dim=1)    # [NB,4,H,W]
    e = GRID - 1
    corners = torch.stack([X[..., 0, 0], Y[..., 0, 0], X[..., 0, e], Y[..., 0, e],
                           X[..., e, 0], Y[..., e, 0], X[..., e, e], Y[..., e, e]], dim=1)  # [NB,8,H,W]
    return loc, corners


def head_forward(class_features, feature_maps, tn, simple_affine, inverse, emulate=False,
                 return_intermediates=False, class_chunk=8):
    """Whole head for normalised class features [C,D,15,15] and image maps [B,D,H,W].
    Returns loc [B,C,4,H,W], score [B,C,1,H,W], corners [B,C,8,H,W].  head.py:308-435."""
    B, D, H, W = feature_maps.shape
    C = class_features.shape[0]
    f = l2_normalize(feature_maps, 1e-5)                   # head.py:339
    locs, scores, corners_all, inter = [], [], [], {}
    for c0 in range(0, C, class_chunk):                    # classes are independent in eval mode
        cf = class_features[c0:c0 + class_chunk]
        Cc = cf.shape[0]
        corr = correlate(cf, f, emulate)                   # [B*Cc,225,H,W]
        p, z = transform_net(corr, tn, emulate)
        theta = theta_from_params(p, simple_affine, inverse)
        score = resample_and_pool(corr, theta, emulate)
        loc, corners = boxes_and_corners(theta, H, W)
        locs.append(loc.view(B, Cc, 4, H, W))
        scores.append(score.view(B, Cc, 1, H, W))
        corners_all.append(corners.view(B, Cc, 8, H, W))
        if return_intermediates:
            inter.setdefault("corr", []).append(corr.view(B, Cc, NUM_CH, H, W))
            inter.setdefault("z", []).append(z.view(B, Cc, NUM_CH, H, W))
            inter.setdefault("params", []).append(p.view(B, Cc, -1, H, W))
    out = (torch.cat(locs, 1), torch.cat(scores, 1), torch.cat(corners_all, 1))
    if return_intermediates:
        return out + ({k: torch.cat(v, 1) for k, v in inter.items()},)
    return out


def random_transform_net(out_dim, seed=0, spread=0.02):
    """Reference-shaped TransformNet state dict with non-identity output: default conv init,
    randomised BN statistics/affine, linear.weight ~ N(0, spread).  (The reference's own init,
    head.py:631-642, regresses the exact identity everywhere and would hide the affine paths.)"""
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

    conv("conv.0", 128, NUM_CH, 7)
    bn("conv.1", 128)
    conv("conv.3", 64, 128, 5)
    bn("conv.4", 64)
    tn["linear.weight"] = spread * torch.randn(out_dim, 64, 5, 5, generator=g)
    bias = torch.zeros(out_dim)
    if out_dim == 6:
        bias[0] = 1
        bias[4] = 1
    else:
        bias[0] = 1
        bias[2] = 1
    tn["linear.bias"] = bias
    return tn
