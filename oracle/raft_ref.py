"""Oracle (test infrastructure): RAFT (basic, non-small) flow in functional torch fp32.

Follows RAFT/raft.py:87-146, RAFT/extractor.py:6-58,118-192, RAFT/update.py:6-136 and
model/modules/flow_comp_raft.py:39-55.  ``sd`` holds the reference's RAFT state_dict
(keys without the DataParallel ``module.`` prefix).
"""
import torch
import torch.nn.functional as F

from . import ops_ref


def _cv(sd, k, x, stride=1, pad=0):
    return F.conv2d(x, sd[k + ".weight"], sd[k + ".bias"], stride=stride, padding=pad)


def _norm(sd, k, x, kind):
    if kind == "instance":                       # nn.InstanceNorm2d default: no affine, no running stats
        return F.instance_norm(x, eps=1e-5)
    return F.batch_norm(x, sd[k + ".running_mean"], sd[k + ".running_var"], sd[k + ".weight"],
                        sd[k + ".bias"], training=False, eps=1e-5)


def _res_block(sd, p, x, kind, stride):
    """extractor.py:6-58 (ResidualBlock)."""
    y = F.relu(_norm(sd, p + ".norm1", _cv(sd, p + ".conv1", x, stride, 1), kind))
    y = F.relu(_norm(sd, p + ".norm2", _cv(sd, p + ".conv2", y, 1, 1), kind))
    if stride != 1:
        x = _norm(sd, p + ".norm3", _cv(sd, p + ".downsample.0", x, stride, 0), kind)
    return F.relu(x + y)


def encoder(sd, p, x, kind):
    """extractor.py:168-192 (BasicEncoder.forward, eval)."""
    x = F.relu(_norm(sd, p + ".norm1", _cv(sd, p + ".conv1", x, 2, 3), kind))
    for li, stride in ((1, 1), (2, 2), (3, 2)):
        x = _res_block(sd, f"{p}.layer{li}.0", x, kind, stride)
        x = _res_block(sd, f"{p}.layer{li}.1", x, kind, 1)
    return _cv(sd, p + ".conv2", x)


def update_block(sd, net, inp, corr, flow):
    """update.py:89-97 (motion encoder), :45-60 (SepConvGRU), :13-14,:131-136 (heads)."""
    u = "update_block."
    cor = F.relu(_cv(sd, u + "encoder.convc1", corr))
    cor = F.relu(_cv(sd, u + "encoder.convc2", cor, 1, 1))
    flo = F.relu(_cv(sd, u + "encoder.convf1", flow, 1, 3))
    flo = F.relu(_cv(sd, u + "encoder.convf2", flo, 1, 1))
    mot = F.relu(_cv(sd, u + "encoder.conv", torch.cat([cor, flo], 1), 1, 1))
    x = torch.cat([inp, mot, flow], 1)
    for tag, pad in (("1", (0, 2)), ("2", (2, 0))):
        hx = torch.cat([net, x], 1)
        z = torch.sigmoid(_cv(sd, u + "gru.convz" + tag, hx, 1, pad))
        r = torch.sigmoid(_cv(sd, u + "gru.convr" + tag, hx, 1, pad))
        q = torch.tanh(_cv(sd, u + "gru.convq" + tag, torch.cat([r * net, x], 1), 1, pad))
        net = (1 - z) * net + z * q
    dflow = _cv(sd, u + "flow_head.conv2", F.relu(_cv(sd, u + "flow_head.conv1", net, 1, 1)), 1, 1)
    up_mask = 0.25 * _cv(sd, u + "mask.2", F.relu(_cv(sd, u + "mask.0", net, 1, 1)))
    return net, up_mask, dflow


def raft_forward(sd, image1, image2, iters=20, return_lowres=False):
    """raft.py:87-146 with test_mode=True; returns the upsampled flow of the last iteration."""
    image1, image2 = image1.contiguous(), image2.contiguous()
    n = image1.shape[0]
    fm = encoder(sd, "fnet", torch.cat([image1, image2], 0), "instance").float()
    pyr = ops_ref.corr_pyramid(fm[:n], fm[n:])
    cn = encoder(sd, "cnet", image1, "batch")
    net, inp = torch.tanh(cn[:, :128]), torch.relu(cn[:, 128:])
    N, _, H, W = image1.shape
    h, w = H // 8, W // 8
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    c0 = torch.stack([xs, ys], 0).float()[None].repeat(N, 1, 1, 1).to(image1.device)
    c1 = c0.clone()
    up = None
    for _ in range(iters):
        corr = ops_ref.corr_lookup(pyr, c1)
        net, up_mask, d = update_block(sd, net, inp, corr, c1 - c0)
        c1 = c1 + d
        up = ops_ref.convex_upsample(c1 - c0, up_mask)
    if return_lowres:
        return c1 - c0, up
    return up


def raft_bi(sd, frames, iters=20):
    """flow_comp_raft.py:39-55.  frames [b,l,3,h,w] -> (fwd, bwd) each [b,l-1,2,h,w]."""
    b, l, c, h, w = frames.shape
    a = frames[:, :-1].reshape(-1, c, h, w)
    bb = frames[:, 1:].reshape(-1, c, h, w)
    fw = raft_forward(sd, a, bb, iters)
    bw = raft_forward(sd, bb, a, iters)
    return fw.view(b, l - 1, 2, h, w), bw.view(b, l - 1, 2, h, w)
