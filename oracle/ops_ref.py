"""Oracle (test infrastructure): primitive ops of the ProPainter hot path, plain torch fp32.

Third-party arithmetic restated here (binary-only in the reference's dependency set,
requirements.txt:10-11 pins only lower bounds torch>=1.7.1 / torchvision>=0.8.2):
  * torchvision.ops.deform_conv2d  (call sites model/propainter.py:67,
    model/recurrent_flow_completion.py:42) -> ``deform_conv3x3`` below, written from the
    published modulated-deformable-convolution definition (DCNv2) with torchvision's
    channel conventions: offset channel g*2K+2k+{0:dy,1:dx}, mask
    channel g*K+k, zero padding outside the image.
  * F.grid_sample is used directly (plain PyTorch is the allowed fp32 reference for a
    floating-point kernel).
"""
import torch
import torch.nn.functional as F


def flow_warp(x, flow_hw2, mode="bilinear"):
    """model/modules/flow_loss_utils.py:6-45.  x [n,c,h,w]; flow_hw2 [n,h,w,2] = (dx,dy) in pixels."""
    n, c, h, w = x.shape
    if tuple(flow_hw2.shape[1:3]) != (h, w):
        raise ValueError("spatial sizes of input and flow differ")
    ys, xs = torch.meshgrid(torch.arange(h, device=x.device), torch.arange(w, device=x.device), indexing="ij")
    base = torch.stack((xs, ys), 2).type_as(x)
    g = base + flow_hw2
    gx = 2.0 * g[..., 0] / max(w - 1, 1) - 1.0
    gy = 2.0 * g[..., 1] / max(h - 1, 1) - 1.0
    return F.grid_sample(x, torch.stack((gx, gy), 3), mode=mode, padding_mode="zeros", align_corners=True)


def fb_consistency(flow_fw, flow_bw, a1=0.01, a2=0.5):
    """model/propainter.py:22-31.  flows [n,2,h,w] -> validity mask [n,1,h,w] in {0,1}."""
    bw_warped = flow_warp(flow_bw, flow_fw.permute(0, 2, 3, 1))
    diff = flow_fw + bw_warped
    mag = (flow_fw ** 2).sum(1, keepdim=True) + (bw_warped ** 2).sum(1, keepdim=True)
    return ((diff ** 2).sum(1, keepdim=True) < a1 * mag + a2).to(flow_fw)


def deform_conv3x3(x, offset, mask, weight, bias):
    """Modulated deformable 3x3 conv, stride 1, pad 1, dil 1, groups 1 (torchvision.ops.deform_conv2d).

    x [B,C,H,W]; offset [B,G*18,H,W]; mask [B,G*9,H,W]; weight [Co,C,3,3]; bias [Co].
    """
    B, C, H, W = x.shape
    K = 9
    G = offset.shape[1] // (2 * K)
    cg = C // G
    dev, dt = x.device, x.dtype
    off = offset.view(B, G, K, 2, H, W)
    ky = torch.arange(3, device=dev, dtype=dt).repeat_interleave(3).view(1, 1, K, 1, 1)
    kx = torch.arange(3, device=dev, dtype=dt).repeat(3).view(1, 1, K, 1, 1)
    yy = torch.arange(H, device=dev, dtype=dt).view(1, 1, 1, H, 1)
    xx = torch.arange(W, device=dev, dtype=dt).view(1, 1, 1, 1, W)
    py = yy - 1 + ky + off[:, :, :, 0]          # [B,G,K,H,W]
    px = xx - 1 + kx + off[:, :, :, 1]
    inside = (py > -1) & (py < H) & (px > -1) & (px < W)
    y0 = torch.floor(py)
    x0 = torch.floor(px)
    ly, lx = py - y0, px - x0
    hy, hx = 1 - ly, 1 - lx
    xf = x.reshape(B, G, cg, H * W)
    out = torch.zeros(B, G, cg, K, H, W, device=dev, dtype=dt)
    for dy, dx, wgt in ((0, 0, hy * hx), (0, 1, hy * lx), (1, 0, ly * hx), (1, 1, ly * lx)):
        yi = (y0 + dy).long()
        xi = (x0 + dx).long()
        ok = inside & (yi >= 0) & (yi <= H - 1) & (xi >= 0) & (xi <= W - 1)
        lin = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).view(B, G, 1, K * H * W).expand(B, G, cg, K * H * W)
        v = torch.gather(xf, 3, lin).view(B, G, cg, K, H, W)
        out = out + v * (wgt * ok.to(dt)).unsqueeze(2)
    out = out * mask.view(B, G, 1, K, H, W)
    cols = out.reshape(B, C * K, H * W)                       # (c, k) ordering == weight.view(Co, C*9)
    res = torch.matmul(weight.reshape(weight.shape[0], C * K), cols)
    return res.view(B, -1, H, W) + bias.view(1, -1, 1, 1)


def corr_pyramid(f1, f2, levels=4):
    """RAFT/corr.py:13-27,52-60.  f1,f2 [B,D,h,w] -> list of [B*h*w,1,h/2^i,w/2^i]."""
    B, D, h, w = f1.shape
    c = torch.matmul(f1.view(B, D, h * w).transpose(1, 2), f2.view(B, D, h * w))
    c = c.view(B * h * w, 1, h, w) / torch.sqrt(torch.tensor(float(D)))
    pyr = [c]
    for _ in range(levels - 1):
        c = F.avg_pool2d(c, 2, stride=2)
        pyr.append(c)
    return pyr


def corr_lookup(pyr, coords, radius=4):
    """RAFT/corr.py:29-50 + RAFT/utils/utils.py:57-71.  coords [B,2,h,w] (x,y) -> [B,4*(2r+1)^2,h,w]."""
    B, _, h, w = coords.shape
    r = radius
    cc = coords.permute(0, 2, 3, 1).reshape(B * h * w, 1, 1, 2)
    d = torch.linspace(-r, r, 2 * r + 1, device=coords.device)
    # NB (reference quirk, corr.py:38-44): delta = stack(meshgrid(dy, dx)) is added to (x, y)
    # coords, so the *first* window axis moves x and the second moves y.
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, 2 * r + 1, 2 * r + 1, 2)
    outs = []
    for i, c in enumerate(pyr):
        H, W = c.shape[-2:]
        pos = cc / 2 ** i + delta
        gx = 2 * pos[..., 0:1] / (W - 1) - 1
        gy = 2 * pos[..., 1:2] / (H - 1) - 1
        s = F.grid_sample(c, torch.cat([gx, gy], -1), align_corners=True)
        outs.append(s.view(B, h, w, -1))
    return torch.cat(outs, -1).permute(0, 3, 1, 2).contiguous().float()


def convex_upsample(flow, mask):
    """RAFT/raft.py:73-84.  flow [N,2,h,w], mask [N,576,h,w] -> [N,2,8h,8w]."""
    N, _, h, w = flow.shape
    m = torch.softmax(mask.view(N, 1, 9, 8, 8, h, w), dim=2)
    nb = F.unfold(8 * flow, [3, 3], padding=1).view(N, 2, 9, 1, 1, h, w)
    up = (m * nb).sum(2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(N, 2, 8 * h, 8 * w)


def psnr_u8(a, b):
    """core/metrics.py:20-36 (calculate_psnr): 20*log10(255/sqrt(mse)) over float64, inf if identical."""
    a = torch.as_tensor(a).to(torch.float64)
    b = torch.as_tensor(b).to(torch.float64)
    mse = ((a - b) ** 2).mean().item()
    if mse == 0:
        return float("inf")
    import math
    return 20.0 * math.log10(255.0 / math.sqrt(mse))
