"""Oracle (test infrastructure): recurrent flow completion net, functional torch fp32.

Follows model/recurrent_flow_completion.py: P3DBlock :148-169, SecondOrderDeformableAlignment
:9-44, BidirectionalPropagation :46-124, RecurrentFlowCompleteNet.forward :272-309,
forward_bidirect_flow :312-337, combine_flow :340-347 (eval mode: no edge branch).
"""
import torch
import torch.nn.functional as F

from .ops_ref import deform_conv3x3


def _lrelu(x, s):
    return F.leaky_relu(x, s)


def _c3(sd, k, x, stride=1, pad=0, dil=1):
    return F.conv3d(x, sd[k + ".weight"], sd[k + ".bias"], stride=stride, padding=pad, dilation=dil)


def _c2(sd, k, x, pad=1):
    return F.conv2d(x, sd[k + ".weight"], sd[k + ".bias"], padding=pad)


def _p3d(sd, p, x, stride):
    """:148-169 (use_residual=0 everywhere in this net)."""
    y = _lrelu(_c3(sd, p + ".conv1.0", x, (1, stride, stride), (0, 1, 1)), 0.2)
    return _c3(sd, p + ".conv2.0", y, 1, (2, 0, 0), (2, 1, 1))


def _up2_conv(sd, k, x):
    """deconv :127-146: bilinear x2 (align_corners=True) then 3x3 conv."""
    return _c2(sd, k + ".conv", F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True))


def second_order_align(sd, p, feat2, cond):
    """:30-44.  feat2 [b,256,h,w], cond [b,384,h,w]."""
    o = cond
    for i in (0, 2, 4):
        o = _lrelu(_c2(sd, f"{p}.conv_offset.{i}", o), 0.1)
    o = _c2(sd, p + ".conv_offset.6", o)
    o1, o2, m = torch.chunk(o, 3, dim=1)
    offset = 5.0 * torch.tanh(torch.cat((o1, o2), 1))
    return deform_conv3x3(feat2, offset, torch.sigmoid(m), sd[p + ".weight"], sd[p + ".bias"])


def propagate(sd, p, x):
    """:67-124.  x [b,t,128,h,w] -> same."""
    b, t, c, h, w = x.shape
    spatial = [x[:, i] for i in range(t)]
    done = {}
    for name in ("backward_", "forward_"):
        order = list(range(t))[::-1] if name == "backward_" else list(range(t))
        hist = []
        prop = x.new_zeros(b, c, h, w)
        for i, idx in enumerate(order):
            cur = spatial[idx]
            if i > 0:
                n2 = hist[-2] if i > 1 else torch.zeros_like(prop)
                cond = torch.cat([prop, cur, n2], 1)
                prop = second_order_align(sd, f"{p}.deform_align.{name}", torch.cat([prop, n2], 1), cond)
            parts = [cur] + [done[k][idx] for k in done] + [prop]
            f = torch.cat(parts, 1)
            y = _c2(sd, f"{p}.backbone.{name}.2", _lrelu(_c2(sd, f"{p}.backbone.{name}.0", f), 0.1))
            prop = prop + y
            hist.append(prop)
        done[name] = hist[::-1] if name == "backward_" else hist
    outs = [F.conv2d(torch.cat([done["backward_"][i], done["forward_"][i]], 1),
                     sd[p + ".fusion.weight"], sd[p + ".fusion.bias"]) for i in range(t)]
    return torch.stack(outs, 1) + x


def rfc_forward(sd, masked_flows, masks):
    """:272-309.  masked_flows [b,t,2,h,w], masks [b,t,1,h,w] -> flow [b,t,2,h,w]."""
    b, t, _, h, w = masked_flows.shape
    inp = torch.cat((masked_flows.permute(0, 2, 1, 3, 4), masks.permute(0, 2, 1, 3, 4)), 1)
    x = F.pad(inp, (2, 2, 2, 2, 0, 0), mode="replicate")
    x = _lrelu(_c3(sd, "downsample.0", x, (1, 2, 2)), 0.2)
    e1 = _lrelu(_p3d(sd, "encoder1.0", x, 1), 0.2)
    e1 = _lrelu(_p3d(sd, "encoder1.2", e1, 2), 0.2)
    e2 = _lrelu(_p3d(sd, "encoder2.0", e1, 1), 0.2)
    e2 = _lrelu(_p3d(sd, "encoder2.2", e2, 2), 0.2)
    m = e2
    for i, d in ((0, 3), (2, 2), (4, 1)):
        m = _lrelu(_c3(sd, f"mid_dilation.{i}", m, 1, (0, d, d), (1, d, d)), 0.2)
    fp = propagate(sd, "feat_prop_module", m.permute(0, 2, 1, 3, 4)).reshape(-1, 128, h // 8, w // 8)
    e1f = e1.permute(0, 2, 1, 3, 4).reshape(b * t, e1.shape[1], e1.shape[3], e1.shape[4])
    d2 = _lrelu(_up2_conv(sd, "decoder2.2", _lrelu(_c2(sd, "decoder2.0", fp), 0.2)), 0.2) + e1f
    d1 = _lrelu(_up2_conv(sd, "decoder1.2", _lrelu(_c2(sd, "decoder1.0", d2), 0.2)), 0.2)
    fl = _up2_conv(sd, "upsample.2", _lrelu(_c2(sd, "upsample.0", d1), 0.2))
    return fl.view(b, t, 2, h, w)


def forward_bidirect_flow(sd, flows_bi, masks):
    """:312-337 (eval).  flows_bi (f,b) each [b,t-1,2,h,w]; masks [b,t,1,h,w]."""
    mf, mb = masks[:, :-1].contiguous(), masks[:, 1:].contiguous()
    pf = rfc_forward(sd, flows_bi[0] * (1 - mf), mf)
    pb = rfc_forward(sd, torch.flip(flows_bi[1] * (1 - mb), dims=[1]), torch.flip(mb, dims=[1]))
    return pf, torch.flip(pb, dims=[1])


def combine_flow(flows_bi, pred_bi, masks):
    """:340-347."""
    mf, mb = masks[:, :-1].contiguous(), masks[:, 1:].contiguous()
    return pred_bi[0] * mf + flows_bi[0] * (1 - mf), pred_bi[1] * mb + flows_bi[1] * (1 - mb)
