"""CPU tests of the kernels' per-element rules (propainter_b200/csrc/pp_elem.cuh compiled for the host by
tests/hostsim) against the oracle.  These pin index arithmetic, padding rules, rounding modes and
threshold placement before any GPU time is spent; the GPU parity tests proper live in test_gpu_*.py."""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

from oracle import generator_ref, ops_ref

FP = ctypes.POINTER(ctypes.c_float)


def fp(t):
    assert t.dtype == torch.float32 and t.is_contiguous()
    return ctypes.cast(t.data_ptr(), FP)


def smooth_flow(gen, n, H, W, amp=4.0):
    z = torch.randn(n, 2, H // 8 + 2, W // 8 + 2, generator=gen) * amp
    return F.interpolate(z, size=(H, W), mode="bicubic", align_corners=False).contiguous()


def test_img_prop_scan(hostsim):
    gen = torch.Generator().manual_seed(0)
    T, H, W = 6, 40, 56
    frames = (torch.rand(1, T, 3, H, W, generator=gen) * 2 - 1)
    ff = smooth_flow(gen, T - 1, H, W).view(1, T - 1, 2, H, W)
    fb = (-ff + 0.3 * smooth_flow(gen, T - 1, H, W).view(1, T - 1, 2, H, W)).contiguous()
    masks = torch.zeros(1, T, 1, H, W)
    masks[..., 10:30, 15:40] = 1
    masked = (frames * (1 - masks)).contiguous()
    for nearest in (1, 0):
        ref_f, ref_m = generator_ref.img_propagation(masked, ff, fb, masks, "nearest" if nearest else "bilinear")
        of, om = torch.empty(T, 3, H, W), torch.empty(T, 1, H, W)
        hostsim.hs_img_prop_scan(fp(masked), fp(ff), fp(fb), fp(masks), fp(of), fp(om), T, H, W, nearest)
        assert 0.02 < ref_m.mean() < masks.mean()           # the propagation really fills pixels
        mism = (om != ref_m[0]).float().mean().item()
        assert mism < 2e-3, mism                            # discontinuous rule: allow isolated threshold flips
        bad = ((of - ref_f[0]).abs() > 1e-5).float().mean().item()
        assert bad < 5e-3, bad


def test_prop_cond(hostsim):
    gen = torch.Generator().manual_seed(1)
    h, w, C = 15, 27, 16
    cur = torch.randn(h, w, C, generator=gen)
    prop = torch.randn(h, w, C, generator=gen)
    f1 = smooth_flow(gen, 1, h * 8, w * 8, 3.0)[0, :, ::8, ::8].contiguous()
    f2 = (-f1 + 0.4 * torch.randn(2, h, w, generator=gen)).contiguous()
    m = (torch.rand(h, w, 2, generator=gen) > 0.5).float()
    ld_c, ld_b = 2 * C + 8, 2 * C + 4
    cond, bb = torch.full((h, w, ld_c), 7.0), torch.full((h, w, ld_b), 7.0)
    fpi, fci = f1.permute(1, 2, 0).contiguous(), f2.permute(1, 2, 0).contiguous()
    hostsim.hs_prop_cond(fp(cur), C, fp(prop), C, fp(fpi), fp(fci), fp(m), fp(cond), ld_c, fp(bb), ld_b, h, w, C, 0)
    valid = ops_ref.fb_consistency(f1[None], f2[None])[0, 0]
    warped = ops_ref.flow_warp(prop.permute(2, 0, 1)[None], f1.permute(1, 2, 0)[None])[0].permute(1, 2, 0)
    assert torch.equal(cond[..., :C], cur)
    assert torch.allclose(cond[..., C:2 * C], warped, atol=1e-5)
    assert torch.equal(cond[..., 2 * C:2 * C + 2], fpi)
    assert (cond[..., 2 * C + 2] != valid).float().mean() < 5e-3
    assert torch.equal(cond[..., 2 * C + 3:2 * C + 5], m) and (cond[..., 2 * C + 5:] == 0).all()
    assert torch.equal(bb[..., :C], cur) and torch.equal(bb[..., 2 * C:2 * C + 2], m) and (bb[..., 2 * C + 2:] == 0).all()
    assert (bb[..., C:2 * C] == 7).all()                    # slot left for the aligned feature
    hostsim.hs_prop_cond(fp(cur), C, fp(prop), C, fp(fpi), fp(fci), fp(m), None, ld_c, fp(bb), ld_b, h, w, C, 1)
    assert torch.equal(bb[..., C:2 * C], cur)


def _pyramid_padded(hostsim, pyr, h, w):
    """oracle pyramid -> the kernels' padded-row layout"""
    out, hl, wl = [], h, w
    for lv in pyr:
        ld = hostsim.hs_corr_ld(wl)
        buf = torch.zeros(lv.shape[0], hl, ld)
        buf[:, :, :wl] = lv[:, 0]
        out.append(buf)
        hl, wl = hl // 2, wl // 2
    return out


def test_corr_pool_and_lookup(hostsim):
    gen = torch.Generator().manual_seed(2)
    B, D, h, w = 2, 32, 16, 22
    f1, f2 = torch.randn(B, D, h, w, generator=gen), torch.randn(B, D, h, w, generator=gen)
    pyr = ops_ref.corr_pyramid(f1, f2)
    pad = _pyramid_padded(hostsim, pyr, h, w)
    hl, wl = h, w
    for l in range(1, 4):
        dst = torch.zeros_like(pad[l])
        hostsim.hs_corr_pool(fp(pad[l - 1]), fp(dst), ctypes.c_long(pad[l].shape[0]), hl, pad[l - 1].shape[2], hl // 2,
                             wl // 2, pad[l].shape[2])
        assert torch.equal(dst[:, :, :wl // 2], pad[l][:, :, :wl // 2]), l
        hl, wl = hl // 2, wl // 2
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    coords = torch.stack([xs, ys], 0).float()[None].repeat(B, 1, 1, 1) + torch.randn(B, 2, h, w, generator=gen) * 6
    ref = ops_ref.corr_lookup(pyr, coords)                       # [B,324,h,w]
    out = torch.empty(B * h * w, 324)
    cpix = coords.permute(0, 2, 3, 1).contiguous()
    hostsim.hs_corr_lookup(fp(pad[0]), fp(pad[1]), fp(pad[2]), fp(pad[3]), fp(cpix), fp(out), ctypes.c_long(B * h * w), h, w)
    got = out.view(B, h, w, 324).permute(0, 3, 1, 2)
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5), (got - ref).abs().max()


def test_convex_upsample(hostsim):
    gen = torch.Generator().manual_seed(3)
    n, h, w = 2, 6, 9
    flow = torch.randn(n, 2, h, w, generator=gen) * 3
    mask = torch.randn(n, 576, h, w, generator=gen) * 4
    ref = ops_ref.convex_upsample(flow, 0.25 * mask)
    out = torch.empty(n, 2, 8 * h, 8 * w)
    mask_px, flow_px = mask.permute(0, 2, 3, 1).contiguous(), flow.permute(0, 2, 3, 1).contiguous()
    hostsim.hs_convex_upsample(fp(mask_px), 576, ctypes.c_float(0.25), fp(flow_px), fp(out), n, h, w)
    assert torch.allclose(out, ref, atol=1e-5, rtol=1e-5)


def test_flow_downsample4(hostsim):
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(3, 24, 40, generator=gen)
    ref = F.interpolate(x[:, None], scale_factor=0.25, mode="bilinear", align_corners=False)[:, 0] / 4.0
    out = torch.empty(3, 6, 10)
    hostsim.hs_flow_ds4(fp(x), fp(out), 3, 24, 40)
    assert torch.allclose(out, ref, atol=1e-7, rtol=1e-6)


def test_deform_gather_columns(hostsim):
    gen = torch.Generator().manual_seed(5)
    for Cin, use_flow, max_res in ((128, True, 3.0), (256, False, 5.0)):
        H, W, Co = 7, 10, 128
        x = torch.randn(1, Cin, H, W, generator=gen)
        o = torch.randn(1, 432, H, W, generator=gen) * 1.5
        flow = torch.randn(1, 2, H, W, generator=gen) * 2
        wgt = torch.randn(Co, Cin, 3, 3, generator=gen) / (Cin * 9) ** 0.5
        bias = torch.randn(Co, generator=gen)
        o1, o2, m = torch.chunk(o, 3, dim=1)
        offset = max_res * torch.tanh(torch.cat((o1, o2), 1))
        if use_flow:
            offset = offset + flow.flip(1).repeat(1, 144, 1, 1)
        ref = ops_ref.deform_conv3x3(x, offset, torch.sigmoid(m), wgt, bias)
        cols = torch.empty(H * W, 9 * Cin)
        xp = x[0].permute(1, 2, 0).contiguous()
        op = o[0].permute(1, 2, 0).contiguous()
        fl = flow[0].permute(1, 2, 0).contiguous()
        hostsim.hs_deform_cols(fp(xp), Cin, fp(op), 432, fp(fl) if use_flow else None, ctypes.c_float(max_res), fp(cols),
                               H, W, Cin)
        wp = wgt.permute(2, 3, 1, 0).reshape(9 * Cin, Co)          # the product's packed layout: row = tap*Cin + c
        got = (cols @ wp + bias).view(H, W, Co).permute(2, 0, 1)
        assert torch.allclose(got, ref[0], atol=2e-4, rtol=1e-4), (got - ref[0]).abs().max()


def test_ffn_overlap_add(hostsim):
    gen = torch.Generator().manual_seed(6)
    for (h, w) in ((15, 21), (32, 32)):
        frames, CH = 2, 8
        fh, fw = (h - 1) // 3 + 1, (w - 1) // 3 + 1
        n = frames * fh * fw
        Y = torch.randn(n, 49 * CH, generator=gen)                  # reference column order: c*49 + tap
        ones = torch.ones(frames, 49, fh * fw)
        norm = F.fold(ones, (h, w), (7, 7), padding=3, stride=3)
        y = F.fold(Y.view(frames, fh * fw, 49 * CH).permute(0, 2, 1), (h, w), (7, 7), padding=3, stride=3)
        ref = F.gelu(F.unfold(y / norm, (7, 7), padding=3, stride=3).permute(0, 2, 1).reshape(n, 49 * CH))
        perm = torch.arange(49 * CH).view(CH, 49).t().reshape(-1)    # tap-major position -> reference column
        Z = torch.empty(n, 49 * CH)
        Yp = Y[:, perm].contiguous()
        hostsim.hs_ffn_overlap_add(fp(Yp), 49 * CH, fp(Z), 49 * CH, frames, h, w, CH)
        assert torch.allclose(Z, ref[:, perm], atol=1e-5, rtol=1e-5)


def test_u8_and_composite(hostsim):
    from oracle import pipeline_ref
    gen = torch.Generator().manual_seed(7)
    T, H, W = 4, 6, 8
    u8 = torch.randint(0, 256, (T, H, W, 3), generator=gen, dtype=torch.uint8)
    out = torch.empty(T, 3, H, W)
    hostsim.hs_u8_to_frames(ctypes.c_void_p(u8.data_ptr()), fp(out), T, H, W)
    assert torch.equal(out, pipeline_ref.to_float_frames(u8.numpy())[0])
    # compositing: replay the reference's numpy sequence for frames visited twice
    masks = (torch.rand(T, 1, H, W, generator=gen) > 0.5).float()
    comp = torch.zeros(T, H, W, 3, dtype=torch.uint8)
    ref = [None] * T
    for visit, ids in enumerate(([0, 1, 2], [1, 2, 3], [2, 3])):
        pred = torch.rand(len(ids), 3, H, W, generator=gen) * 2 - 1
        first = [int(ref[i] is None) for i in ids]
        pr = ((pred + 1) / 2).permute(0, 2, 3, 1).numpy() * 255
        bm = masks[ids].permute(0, 2, 3, 1).numpy().astype(np.uint8)
        for k, i in enumerate(ids):
            img = np.array(pr[k]).astype(np.uint8) * bm[k] + u8[i].numpy() * (1 - bm[k])
            ref[i] = img if ref[i] is None else (ref[i].astype(np.float32) * 0.5 + img.astype(np.float32) * 0.5)
            ref[i] = ref[i].astype(np.uint8)
        fr = (ctypes.c_int * len(ids))(*ids)
        fs = (ctypes.c_int * len(ids))(*first)
        hostsim.hs_composite(fp(pred), fp(masks), ctypes.c_void_p(u8.data_ptr()), ctypes.c_void_p(comp.data_ptr()),
                             len(ids), fr, fs, H, W)
    assert np.array_equal(comp.numpy(), np.stack(ref, 0))


def test_upsample2x(hostsim):
    gen = torch.Generator().manual_seed(9)
    for (n, h, w, C) in ((2, 15, 27, 8), (1, 30, 54, 4)):
        x = torch.randn(n, C, h, w, generator=gen)
        ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        xp = x.permute(0, 2, 3, 1).contiguous()
        out = torch.empty(n, 2 * h, 2 * w, C)
        hostsim.hs_upsample2x(fp(xp), fp(out), n, h, w, C)
        assert torch.allclose(out.permute(0, 3, 1, 2), ref, atol=1e-6, rtol=1e-6)


def test_mask_dilate(hostsim):
    import scipy.ndimage
    gen = torch.Generator().manual_seed(12)
    T, H, W = 3, 37, 53
    m = (torch.rand(T, H, W, generator=gen) > 0.985).to(torch.uint8) * 255
    m[0, 0, 0] = 255
    m[1, H - 1, W - 1] = 7
    for it in (0, 1, 4, 8):
        ref = np.stack([scipy.ndimage.binary_dilation(m[i].numpy(), iterations=it) if it else m[i].numpy() > 0 for i in range(T)])
        out = torch.empty(T, H, W)
        hostsim.hs_mask_dilate(ctypes.c_void_p(m.data_ptr()), fp(out), T, H, W, it)
        assert np.array_equal(out.numpy() > 0.5, ref), it
