"""RAFT optical flow (basic model) on the B200 hot path.

Drop-in for the reference's ``RAFT`` (RAFT/raft.py:24-146): same constructor argument, same
``forward(image1, image2, iters, flow_init, test_mode)`` result in test mode, same state_dict.
What differs is the execution plan:
  * features live pixel-major (channels-last); convs are cuDNN through torch
  * the all-pairs volume + pyramid (corr.py:13-27) is built by ``ops.corr_build`` (fp32-accurate
    3xTF32 tensor-core GEMM) into row-padded planes; the 4-level 9x9 lookup (corr.py:29-50) is one
    kernel writing the 324-channel pixel-major tensor the motion encoder consumes
  * z and r gates of the SepConvGRU share one conv (concatenated weights); eval BatchNorm is folded
    into the cnet convs; the mask head + convex upsampling run only on the last iteration (the
    reference computes them 20x and keeps one, raft.py:135-146)
  * ``flows_bidirectional`` encodes every frame once for both directions (the reference encodes each
    frame up to 4x, flow_comp_raft.py:48-49); per-sample InstanceNorm makes this exactly equivalent
"""
import torch
import torch.nn.functional as F

from .. import ops
from .._params import ParamNet
from ..nn_util import as_nchw, as_pm, cl, conv
from ..schemas import raft_schema


class RAFT(ParamNet):
    hidden_dim = 128
    context_dim = 128

    def __init__(self, args=None, seed=None):
        if args is not None and getattr(args, "small", False):
            raise NotImplementedError("only the basic RAFT model is on ProPainter's path (flow_comp_raft.py:15)")
        super().__init__(raft_schema(), seed=seed)
        self.args = args

    # ------------------------------------------------------------------ weights
    def _wb(self, key, bn=None):
        """(weight, bias) of a conv, channels_last, with eval-BatchNorm `bn` folded in if given."""
        def build():
            w, b = self.P[key + ".weight"], self.P[key + ".bias"]
            if bn is not None and (bn + ".running_var") in self.P:
                s = self.P[bn + ".weight"] / torch.sqrt(self.P[bn + ".running_var"] + 1e-5)
                w = w * s.view(-1, 1, 1, 1)
                b = (b - self.P[bn + ".running_mean"]) * s + self.P[bn + ".bias"]
            return cl(w), b.contiguous()
        return self.packed("wb:" + key, build)

    def _motion_out(self):
        """update_block.encoder.conv with its 126 outputs padded to 128 (two zero filters): keeps the output
        rows 16-byte aligned; the pad slots are overwritten by the flow channels in raft_pack_motion."""
        def build():
            w, b = self.P["update_block.encoder.conv.weight"], self.P["update_block.encoder.conv.bias"]
            w = torch.cat([w, w.new_zeros(2, *w.shape[1:])], 0)
            return cl(w), torch.cat([b, b.new_zeros(2)]).contiguous()
        return self.packed("motion_out", build)

    def _gru(self, tag):
        """SepConvGRU weights of one pass (update.py:45-60), input channels [h(128) | inp(128) | motion(128)] (:129-130).
        z and r share one conv.  The `inp` (context) channels do not change over the refinement iterations, so their
        share of every gate conv is split off: (w_zr, w_q) act on the per-iteration [h | motion] buffers, (w_zr_inp,
        b_zr) / (w_q_inp, b_q) are convolved with `inp` once per clip and enter through the gate kernels' `pre` term."""
        def build():
            u = "update_block.gru."
            wzr = torch.cat([self.P[u + f"convz{tag}.weight"], self.P[u + f"convr{tag}.weight"]], 0)
            bzr = torch.cat([self.P[u + f"convz{tag}.bias"], self.P[u + f"convr{tag}.bias"]], 0)
            wq, bq = self.P[u + f"convq{tag}.weight"], self.P[u + f"convq{tag}.bias"]
            dyn = lambda w: cl(torch.cat([w[:, :128], w[:, 256:]], 1))
            ctx = lambda w: cl(w[:, 128:256])
            return dyn(wzr), dyn(wq), (ctx(wzr), bzr.contiguous()), (ctx(wq), bq.contiguous())
        return self.packed("gru" + tag, build)

    # ------------------------------------------------------------------ encoders (extractor.py:168-192)
    def _encode(self, p, x):
        """BasicEncoder.forward extractor.py:168-192.  fnet: conv -> InstanceNorm -> ReLU with the norm, the ReLU and the
        block's `relu(x + y)` as one pp_instance_norm call on the raw conv output (the conv bias cancels under the
        per-channel mean subtraction, so it is not even added).  cnet: eval BatchNorm folded into the conv; bias, ReLU
        and the residual add + ReLU are one pp_bias_act pass."""
        inst = p == "fnet"

        def cn(key, bn, t, stride=1, pad=1, relu=True, res=None):
            if inst:
                w, _ = self._wb(key)
                y = as_pm(F.conv2d(t, w, None, stride=stride, padding=pad))
                return as_nchw(ops.instance_norm(y, relu=relu, res=None if res is None else as_pm(res),
                                                 post_relu=res is not None, out=y))
            return conv(t, self._wb(key, bn), stride, pad, act="relu" if relu else "none", res=res, post_relu=res is not None)

        x = cn(p + ".conv1", p + ".norm1", x, 2, 3)
        for li, stride in ((1, 1), (2, 2), (3, 2)):
            for bi in (0, 1):
                q = f"{p}.layer{li}.{bi}"
                s = stride if bi == 0 else 1
                y = cn(q + ".conv1", q + ".norm1", x, s, 1)
                if s != 1:
                    x = cn(q + ".downsample.0", q + ".norm3", x, s, 0, relu=False)
                x = cn(q + ".conv2", q + ".norm2", y, 1, 1, res=x)              # relu(x + relu(norm(conv2(y))))
        return conv(x, self._wb(p + ".conv2"))

    def encode_frames(self, frames):
        """frames [n,3,H,W] -> (fmap pixel-major [n, h*w, 256], net [n,128,h,w], inp [n,128,h,w])."""
        x = frames.contiguous(memory_format=torch.channels_last)
        fmap = as_pm(self._encode("fnet", x).float())
        n, h, w, d = fmap.shape
        c = self._encode("cnet", x)
        net, inp = torch.tanh(c[:, :128]), torch.relu(c[:, 128:])
        return fmap.view(n, h * w, d), net, inp, (h, w)

    # ------------------------------------------------------------------ refinement loop (raft.py:122-146)
    def _refine(self, fmap, idx1, idx2, net, inp, hw, iters, flow_init=None):
        h, w = hw
        B = idx1.numel()
        dev = fmap.device
        levels = ops.corr_alloc(B, h, w, dev)
        ops.corr_build(fmap, idx1, idx2, levels, h, w)
        ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
        c0 = torch.stack([xs, ys], -1).float()[None].expand(B, h, w, 2).contiguous()     # coords_grid (utils.py:74-77)
        c1 = c0.clone()
        if flow_init is not None:
            c1 = c1 + as_pm(flow_init)
        u = "update_block."
        corr = torch.empty(B, h, w, 324, device=dev)
        # persistent GRU buffers (no torch.cat inside the loop):
        #   HX = [net | motion(126) flow(2)]  -> z/r gate convs;   RX = [r*net | motion flow] -> candidate conv
        # the context channels `inp` are iteration-invariant: their conv contributions (+ biases) are computed here, once
        HX = torch.empty(B, h, w, 256, device=dev)
        RX = torch.empty(B, h, w, 256, device=dev)
        HX[..., :128] = as_pm(net)
        pre = {}
        for tag, pad in (("1", (0, 2)), ("2", (2, 0))):
            _, _, zr_ctx, q_ctx = self._gru(tag)
            pre[tag] = (as_pm(conv(inp, zr_ctx, 1, pad)), as_pm(conv(inp, q_ctx, 1, pad)))
        netv, z = HX[..., :128], torch.empty(B, h, w, 128, device=dev)
        netc = torch.empty(B, h, w, 128, device=dev)            # dense copy of the state for the flow / mask heads
        mot_in = torch.empty(B, h, w, 256, device=dev)          # [cor(192) | flo(64)] without a torch.cat (update.py:95)
        mw, mb = self._motion_out()
        for _ in range(iters):
            ops.corr_lookup(levels, c1, corr)
            flow_pm = c1 - c0
            flow = as_nchw(flow_pm)
            cor = conv(as_nchw(corr), self._wb(u + "encoder.convc1"), act="relu")
            conv(cor, self._wb(u + "encoder.convc2"), 1, 1, act="relu", out=as_nchw(mot_in[..., :192]))
            flo = conv(flow, self._wb(u + "encoder.convf1"), 1, 3, act="relu")
            conv(flo, self._wb(u + "encoder.convf2"), 1, 1, act="relu", out=as_nchw(mot_in[..., 192:]))
            mot = F.conv2d(as_nchw(mot_in), mw, None, padding=1)                   # 126 real + 2 pad channels, raw
            ops.raft_pack_motion(as_pm(mot), flow_pm, HX[..., 128:], RX[..., 128:], bias=mb)   # + bias + ReLU (update.py:96)
            for tag, pad in (("1", (0, 2)), ("2", (2, 0))):
                gw, qw, _, _ = self._gru(tag)
                ops.gru_gate(as_pm(F.conv2d(as_nchw(HX), gw, None, padding=pad)), None, netv, z, RX[..., :128], pre=pre[tag][0])
                ops.gru_update(as_pm(F.conv2d(as_nchw(RX), qw, None, padding=pad)), None, z, netv,
                               net_copy=netc if tag == "2" else None, pre=pre[tag][1])
            net = as_nchw(netc)
            d = conv(conv(net, self._wb(u + "flow_head.conv1"), 1, 1, act="relu"), self._wb(u + "flow_head.conv2"), 1, 1)
            c1 = c1 + as_pm(d)
        flow_lr = c1 - c0
        mask = conv(conv(net, self._wb(u + "mask.0"), 1, 1, act="relu"), self._wb(u + "mask.2"))
        up = ops.convex_upsample(as_pm(mask), flow_lr.contiguous(), 0.25)
        return as_nchw(flow_lr), up

    @torch.no_grad()
    def forward(self, image1, image2, iters=12, flow_init=None, test_mode=True):
        """raft.py:87-146.  image1/2 [N,3,H,W] in [-1,1] -> (flow_lowres [N,2,H/8,W/8], flow_up [N,2,H,W])."""
        if not test_mode:
            raise NotImplementedError("training-mode flow_predictions list is outside the inference hot path")
        n = image1.shape[0]
        fmap, net, inp, hw = self.encode_frames(torch.cat([image1, image2], 0))
        idx1 = torch.arange(n, device=image1.device, dtype=torch.int32)
        return self._refine(fmap, idx1, idx1 + n, net[:n], inp[:n], hw, iters, flow_init)

    @torch.no_grad()
    def flows_bidirectional(self, frames, iters=20):
        """frames [l,3,H,W] -> (forward flows i->i+1, backward flows i+1->i), each [l-1,2,H,W].
        Encoders + 20 refinement iterations replay as one CUDA graph per clip shape."""
        return self.graphs(("raft_bi", iters), lambda fr: self._flows_bidirectional(fr, iters), frames.contiguous())

    def _flows_bidirectional(self, frames, iters):
        l = frames.shape[0]
        fmap, net, inp, hw = self.encode_frames(frames)
        a = torch.arange(l - 1, device=frames.device, dtype=torch.int32)
        idx1, idx2 = torch.cat([a, a + 1]), torch.cat([a + 1, a])
        sel = idx1.long()
        _, up = self._refine(fmap, idx1, idx2, net[sel], inp[sel], hw, iters)
        return up[:l - 1], up[l - 1:]
