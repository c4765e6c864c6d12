"""Oracle (test infrastructure): InpaintGenerator (image propagation, encoder, feature
propagation, sparse transformer, decoder) in functional torch fp32.

Follows model/propainter.py:72-190 (BidirectionalPropagation), :34-69 (DeformableAlignment),
:193-253 (Encoder, deconv), :315-372 (img_propagation, forward) and
model/modules/sparse_transformer.py:7-344.  Eval mode only.
"""
import math

import torch
import torch.nn.functional as F

from .ops_ref import deform_conv3x3, fb_consistency, flow_warp

WIN = (5, 9)
POOL = (4, 4)
HEADS = 4
KS, ST, PD = (7, 7), (3, 3), (3, 3)


def _c2(sd, k, x, stride=1, pad=1, groups=1):
    return F.conv2d(x, sd[k + ".weight"], sd[k + ".bias"], stride=stride, padding=pad, groups=groups)


def _bin(m, th=0.1):
    return (m > th).to(m)


# ---------------------------------------------------------------- propagation
def img_propagation(frames, flows_f, flows_b, masks, mode="nearest"):
    """propainter.py:104-190 with learnable=False (called from :315-317).

    frames [b,t,3,h,w], flows [b,t-1,2,h,w], masks [b,t,1,h,w] -> (frames_out, masks_out).
    """
    b, t, c, h, w = frames.shape
    feats = [frames[:, i] for i in range(t)]
    msks = [masks[:, i] for i in range(t)]
    for direction in ("backward", "forward"):
        order = list(range(t))[::-1] if direction == "backward" else list(range(t))
        nf, nm = [None] * t, [None] * t
        fp = mp = None
        for i, idx in enumerate(order):
            cur, mcur = feats[idx], msks[idx]
            if i == 0:
                fp, mp = cur, mcur
            else:
                if direction == "backward":
                    fl_p, fl_c = flows_f[:, idx], flows_b[:, idx]
                else:
                    fl_p, fl_c = flows_b[:, idx - 1], flows_f[:, idx - 1]
                valid = fb_consistency(fl_p, fl_c)
                warped = flow_warp(fp, fl_p.permute(0, 2, 3, 1), mode)
                mwarp = _bin(flow_warp(mp, fl_p.permute(0, 2, 3, 1)))
                use = _bin(mcur * valid * (1 - mwarp))
                fp = use * warped + (1 - use) * cur
                mp = _bin(mcur * (1 - valid * (1 - mwarp)))
            nf[idx], nm[idx] = fp, mp
        feats, msks = nf, nm
    return torch.stack(feats, 1), torch.stack(msks, 1)


def deform_align(sd, p, x, cond, flow):
    """propainter.py:56-69."""
    o = cond
    for i in (0, 2, 4):
        o = F.leaky_relu(_c2(sd, f"{p}.conv_offset.{i}", o), 0.1)
    o = _c2(sd, p + ".conv_offset.6", o)
    o1, o2, m = torch.chunk(o, 3, dim=1)
    offset = 3.0 * torch.tanh(torch.cat((o1, o2), 1))
    offset = offset + flow.flip(1).repeat(1, offset.shape[1] // 2, 1, 1)
    return deform_conv3x3(x, offset, torch.sigmoid(m), sd[p + ".weight"], sd[p + ".bias"])


def feat_propagation(sd, p, x, flows_f, flows_b, mask, mode="bilinear"):
    """propainter.py:104-190 with learnable=True.  x [b,t,128,h,w], mask [b,t,2,h,w] -> fused [b,t,128,h,w]."""
    b, t, c, h, w = x.shape
    src = [x[:, i] for i in range(t)]
    res = {}
    for name in ("backward_1", "forward_1"):
        order = list(range(t))[::-1] if name == "backward_1" else list(range(t))
        out = [None] * t
        fp = None
        for i, idx in enumerate(order):
            cur, mcur = src[idx], mask[:, idx]
            if i == 0:
                fp = cur
            else:
                if name == "backward_1":
                    fl_p, fl_c = flows_f[:, idx], flows_b[:, idx]
                else:
                    fl_p, fl_c = flows_b[:, idx - 1], flows_f[:, idx - 1]
                valid = fb_consistency(fl_p, fl_c)
                warped = flow_warp(fp, fl_p.permute(0, 2, 3, 1), mode)
                cond = torch.cat([cur, warped, fl_p, valid, mcur], 1)
                fp = deform_align(sd, f"{p}.deform_align.{name}", fp, cond, fl_p)
            z = torch.cat([cur, fp, mcur], 1)
            z = _c2(sd, f"{p}.backbone.{name}.2", F.leaky_relu(_c2(sd, f"{p}.backbone.{name}.0", z), 0.2))
            fp = fp + z
            out[idx] = fp
        res[name] = out
        src = out                      # the forward scan consumes the backward scan's features (:138)
    ob = torch.stack(res["backward_1"], 1).view(-1, c, h, w)
    of = torch.stack(res["forward_1"], 1).view(-1, c, h, w)
    z = torch.cat([ob, of, mask.reshape(-1, 2, h, w)], 1)
    z = _c2(sd, p + ".fuse.2", F.leaky_relu(_c2(sd, p + ".fuse.0", z), 0.2))
    return (z + x.reshape(-1, c, h, w)).view(b, t, c, h, w)


# ---------------------------------------------------------------- conv trunk
def encoder(sd, x):
    """propainter.py:218-232: grouped skip-concat of the layer-8 input into layers 10..16."""
    groups = {10: 2, 12: 4, 14: 8, 16: 1}
    strides = {0: 2, 4: 2}
    out = x
    x0 = None
    for i in range(0, 18, 2):
        if i == 8:
            x0 = out
        if i > 8:
            g = groups[i]
            n, _, h, w = out.shape
            out = torch.cat([x0.view(n, g, -1, h, w), out.view(n, g, -1, h, w)], 2).view(n, -1, h, w)
        out = F.leaky_relu(_c2(sd, f"encoder.layers.{i}", out, strides.get(i, 1), 1, groups.get(i, 1)), 0.2)
    return out


def _up2_conv(sd, k, x):
    return _c2(sd, k + ".conv", F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True))


def decoder(sd, x):
    x = F.leaky_relu(_up2_conv(sd, "decoder.0", x), 0.2)
    x = F.leaky_relu(_c2(sd, "decoder.2", x), 0.2)
    x = F.leaky_relu(_up2_conv(sd, "decoder.4", x), 0.2)
    return _c2(sd, "decoder.6", x)


# ---------------------------------------------------------------- transformer
def _token_grid(hw):
    return tuple((hw[i] + 2 * PD[i] - KS[i]) // ST[i] + 1 for i in range(2))


def soft_split(sd, x, b, hw):
    """sparse_transformer.py:19-31."""
    fh, fw = _token_grid(hw)
    f = F.unfold(x, KS, stride=ST, padding=PD).permute(0, 2, 1)
    f = F.linear(f, sd["ss.embedding.weight"], sd["ss.embedding.bias"])
    return f.reshape(b, -1, fh, fw, f.shape[2])


def soft_comp(sd, x, t, hw):
    """sparse_transformer.py:49-61."""
    b = x.shape[0]
    f = F.linear(x.reshape(b, -1, x.shape[-1]), sd["sc.embedding.weight"], sd["sc.embedding.bias"])
    f = f.view(b * t, -1, f.shape[2]).permute(0, 2, 1)
    f = F.fold(f, hw, KS, stride=ST, padding=PD)
    return _c2(sd, "sc.bias_conv", f)


def _windows(x, heads):
    """sparse_transformer.py:104-115 -> [B, nWin, heads, T, wh*ww, C/heads]."""
    B, T, H, W, C = x.shape
    x = x.view(B, T, H // WIN[0], WIN[0], W // WIN[1], WIN[1], heads, C // heads)
    x = x.permute(0, 2, 4, 6, 1, 3, 5, 7).contiguous()
    return x.view(B, -1, heads, T, WIN[0] * WIN[1], C // heads)


def _rolled_valid_index():
    """sparse_transformer.py:140-153: which of the 4x45 rolled tokens lie outside the own window."""
    e = tuple((i + 1) // 2 for i in WIN)
    ms = []
    for top, left in ((True, True), (True, False), (False, True), (False, False)):
        m = torch.ones(WIN)
        rows = slice(None, -e[0]) if top else slice(e[0], None)
        cols = slice(None, -e[1]) if left else slice(e[1], None)
        m[rows, cols] = 0
        ms.append(m)
    return torch.stack(ms, 0).flatten().nonzero(as_tuple=False).view(-1)


def window_attention(sd, p, x, mask, t_ind):
    """sparse_transformer.py:158-281.  x [b,t,h,w,c]; mask [b,l_t,h,w,1]; t_ind 1-D LongTensor."""
    b, t, h, w, c = x.shape
    wh, ww = WIN
    ch = c // HEADS
    nwh, nww = math.ceil(h / wh), math.ceil(w / ww)
    H2, W2 = nwh * wh, nww * ww
    if H2 > h or W2 > w:
        x = F.pad(x, (0, 0, 0, W2 - w, 0, H2 - h))
        mask = F.pad(mask, (0, 0, 0, W2 - w, 0, H2 - h))
    lin = lambda n, z: F.linear(z, sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"])
    q, k, v = lin("query", x), lin("key", x), lin("value", x)
    wq, wk, wv = _windows(q, HEADS), _windows(k, HEADS), _windows(v, HEADS)
    e = tuple((i + 1) // 2 for i in WIN)
    vi = _rolled_valid_index().to(x.device)
    rk, rv = [], []
    for sy, sx in ((-e[0], -e[1]), (-e[0], e[1]), (e[0], -e[1]), (e[0], e[1])):
        rk.append(_windows(torch.roll(k, (sy, sx), (2, 3)), HEADS))
        rv.append(_windows(torch.roll(v, (sy, sx), (2, 3)), HEADS))
    wk = torch.cat((wk, torch.cat(rk, 4)[:, :, :, :, vi]), 4)
    wv = torch.cat((wv, torch.cat(rv, 4)[:, :, :, :, vi]), 4)
    px = F.conv2d(x.view(b * t, H2, W2, c).permute(0, 3, 1, 2), sd[p + ".pool_layer.weight"],
                  sd[p + ".pool_layer.bias"], stride=POOL, groups=c)
    ph, pw = px.shape[-2:]
    px = px.permute(0, 2, 3, 1).view(b, t, ph * pw, c)
    pk = lin("key", px).view(b, 1, t, ph * pw, HEADS, ch).permute(0, 1, 4, 2, 3, 5).expand(-1, nwh * nww, -1, -1, -1, -1)
    pv = lin("value", px).view(b, 1, t, ph * pw, HEADS, ch).permute(0, 1, 4, 2, 3, 5).expand(-1, nwh * nww, -1, -1, -1, -1)
    wk = torch.cat((wk, pk), 4)
    wv = torch.cat((wv, pv), 4)
    out = torch.zeros_like(wq)
    lt = mask.shape[1]
    wm = F.max_pool2d(mask.view(b * lt, 1, H2, W2), WIN, WIN).view(b, lt, nwh * nww).sum(1)
    scale = 1.0 / math.sqrt(ch)
    for i in range(b):
        mi = wm[i].nonzero(as_tuple=False).view(-1)
        if len(mi) > 0:
            qt = wq[i, mi].reshape(len(mi), HEADS, t * wh * ww, ch)
            kt = wk[i, mi][:, :, t_ind].reshape(len(mi), HEADS, -1, ch)
            vt = wv[i, mi][:, :, t_ind].reshape(len(mi), HEADS, -1, ch)
            a = F.softmax((qt @ kt.transpose(-2, -1)) * scale, dim=-1)
            out[i, mi] = (a @ vt).view(-1, HEADS, t, wh * ww, ch)
        ui = (wm[i] == 0).nonzero(as_tuple=False).view(-1)
        qs, ks, vs = wq[i, ui], wk[i, ui, :, :, :wh * ww], wv[i, ui, :, :, :wh * ww]
        a = F.softmax((qs @ ks.transpose(-2, -1)) * scale, dim=-1)
        out[i, ui] = a @ vs
    out = out.view(b, nwh, nww, HEADS, t, wh, ww, ch).permute(0, 4, 1, 5, 2, 6, 3, 7).reshape(b, t, H2, W2, c)
    out = out[:, :, :h, :w]
    return lin("proj", out)


def fusion_ffn(sd, p, x, hw):
    """sparse_transformer.py:74-101.  x [b,n,512]."""
    fh, fw = _token_grid(hw)
    nv = fh * fw
    x = F.linear(x, sd[p + ".fc1.0.weight"], sd[p + ".fc1.0.bias"])
    b, n, c = x.shape
    ones = x.new_ones(b * n // nv, KS[0] * KS[1], nv)
    norm = F.fold(ones, hw, KS, padding=PD, stride=ST)
    y = F.fold(x.view(-1, nv, c).permute(0, 2, 1), hw, KS, padding=PD, stride=ST)
    y = F.unfold(y / norm, KS, padding=PD, stride=ST).permute(0, 2, 1).contiguous().view(b, n, c)
    return F.linear(F.gelu(y), sd[p + ".fc2.1.weight"], sd[p + ".fc2.1.bias"])


def transformer(sd, x, hw, mask, t_dilation=2, depths=8):
    """sparse_transformer.py:294-344."""
    B, T, H, W, C = x.shape
    assert depths % t_dilation == 0
    sched = [torch.arange(i, T, t_dilation) for i in range(t_dilation)] * (depths // t_dilation)
    for i in range(depths):
        p = f"transformers.transformer.{i}"
        y = F.layer_norm(x, (C,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
        x = x + window_attention(sd, p + ".attention", y, mask, sched[i].to(x.device))
        y = F.layer_norm(x, (C,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"])
        x = x + fusion_ffn(sd, p + ".mlp", y.view(B, T * H * W, C), hw).view(B, T, H, W, C)
    return x


# ---------------------------------------------------------------- generator
def generator_forward(sd, frames, flows, masks_in, masks_upd, l_t, mode="bilinear", t_dilation=2,
                      return_parts=False):
    """propainter.py:319-372 (eval).  frames [b,t,3,H,W] -> [b,l_t,3,H,W]."""
    b, t, _, H, W = frames.shape
    enc = encoder(sd, torch.cat([frames.view(b * t, 3, H, W), masks_in.view(b * t, 1, H, W),
                                 masks_upd.view(b * t, 1, H, W)], 1))
    _, c, h, w = enc.shape
    enc = enc.view(b, t, c, h, w)
    local, ref = enc[:, :l_t], enc[:, l_t:]
    dsf = F.interpolate(flows[0].reshape(-1, 2, H, W), scale_factor=0.25, mode="bilinear",
                        align_corners=False).view(b, l_t - 1, 2, h, w) / 4.0
    dsb = F.interpolate(flows[1].reshape(-1, 2, H, W), scale_factor=0.25, mode="bilinear",
                        align_corners=False).view(b, l_t - 1, 2, h, w) / 4.0
    dm_in = F.interpolate(masks_in.reshape(-1, 1, H, W), scale_factor=0.25, mode="nearest").view(b, t, 1, h, w)
    dm_in_l = dm_in[:, :l_t]
    dm_up_l = F.interpolate(masks_upd[:, :l_t].reshape(-1, 1, H, W), scale_factor=0.25,
                            mode="nearest").view(b, l_t, 1, h, w)
    mp = F.max_pool2d(dm_in_l.reshape(-1, 1, h, w), KS, ST, PD)
    mp = mp.view(b, l_t, 1, mp.shape[-2], mp.shape[-1]).permute(0, 1, 3, 4, 2).contiguous()
    pmask = torch.cat([dm_in_l, dm_up_l], 2)
    local = feat_propagation(sd, "feat_prop_module", local, dsf, dsb, pmask, mode)
    enc = torch.cat((local, ref), 1)
    tok = soft_split(sd, enc.reshape(-1, c, h, w), b, (h, w))
    tok2 = transformer(sd, tok, (h, w), mp, t_dilation)
    tr = soft_comp(sd, tok2, t, (h, w)).view(b, t, -1, h, w)
    enc2 = enc + tr
    out = torch.tanh(decoder(sd, enc2[:, :l_t].reshape(-1, c, h, w))).view(b, l_t, 3, H, W)
    if return_parts:
        return out, {"prop_feat": local, "tokens_in": tok, "tokens_out": tok2, "enc_out": enc2}
    return out
