"""Seeded synthetic clips for parity tests and bench.py (C1 square mask, C2 moving
ellipse, C3 border mask).  Pure torch on CPU; no datasets exist in the build environment."""
import math

import numpy as np
import torch
import torch.nn.functional as F


def smooth_field(gen, c, h, w, cell=12):
    """Low-pass random texture in [0,1] (so RAFT sees trackable structure)."""
    gh, gw = h // cell + 3, w // cell + 3
    z = torch.rand(1, c, gh, gw, generator=gen)
    z = F.interpolate(z, size=(gh * cell, gw * cell), mode="bicubic", align_corners=False)
    fine = torch.rand(1, c, gh * cell // 3, gw * cell // 3, generator=gen)
    z = 0.75 * z + 0.25 * F.interpolate(fine, size=z.shape[-2:], mode="bilinear", align_corners=False)
    return z[0].clamp(0, 1)


def make_clip(T, H, W, mask="ellipse", seed=0, motion=(2.0, 1.0)):
    """Returns frames_u8 [T,H,W,3] uint8 (numpy), flow_masks / masks_dilated [1,T,1,H,W] float {0,1}."""
    gen = torch.Generator().manual_seed(seed)
    mx, my = motion
    pad_x, pad_y = int(abs(mx) * T) + 8, int(abs(my) * T) + 8
    canvas = smooth_field(gen, 3, H + 2 * pad_y + 24, W + 2 * pad_x + 24)
    frames = []
    for t in range(T):
        ox, oy = pad_x + int(round(mx * t)), pad_y + int(round(my * t))
        frames.append(canvas[:, oy:oy + H, ox:ox + W])
    fr = (torch.stack(frames, 0) * 255).round().clamp(0, 255).to(torch.uint8)
    frames_u8 = fr.permute(0, 2, 3, 1).contiguous().numpy()
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    masks = torch.zeros(T, 1, H, W)
    for t in range(T):
        if mask == "square":
            masks[t, 0, H // 4:3 * H // 4, W // 4:3 * W // 4] = 1
        elif mask == "border":
            bh, bw = max(1, round(H * 0.067)), max(1, round(W * 0.067))     # outer ring ~ 25 % of the area
            masks[t, 0] = 1
            masks[t, 0, bh:H - bh, bw:W - bw] = 0
        else:
            cx = W * 0.25 + (2.0 * t) % (W * 0.5)
            cy = H * 0.5 + H * 0.12 * math.sin(t / 7.0)
            ax, ay = W * 0.16, H * 0.16                                    # ~8 % of the area
            masks[t, 0] = ((((xs - cx) / ax) ** 2 + ((ys - cy) / ay) ** 2) <= 1).float()
    masks = masks.unsqueeze(0)
    return frames_u8, masks.clone(), masks.clone()


def dilate_cross(masks, iterations=4):
    """scipy.ndimage.binary_dilation default structure (3x3 cross), iterated (inference_propainter.py:96,105)."""
    m = masks.reshape(-1, 1, *masks.shape[-2:])
    k = torch.tensor([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=m.dtype).view(1, 1, 3, 3)
    for _ in range(iterations):
        m = (F.conv2d(m, k, padding=1) > 0).to(m.dtype)
    return m.view(masks.shape)
