"""Soft split / composition and the mask-guided sparse temporal transformer (B200 execution plan).

Reference: model/modules/sparse_transformer.py (SoftSplit :7-31, SoftComp :34-61, FusionFeedForward
:64-101, SparseWindowAttention :117-281, TemporalSparseTransformer(Block) :284-344).  Parameters live in
the owning ``InpaintGenerator`` (a ParamNet); these helpers only execute.

  * SoftSplit  = unfold(7,3,3) + Linear(6272->512) == one strided 7x7 conv  (no 325 MB im2col buffer)
  * SoftComp   = Linear(512->6272) + fold == one transposed conv + a constant (folded-bias) map
  * attention  = LayerNorm, fused QKV GEMM, pooled K/V, then ``ops.sparse_window_attn`` which gathers
                 own / rolled / pooled keys arithmetically (no roll / window_partition / cat / index copies,
                 no score matrix in HBM) and handles masked and unmasked windows in one call
  * fusion FFN = GEMM, ``ops.ffn_overlap_add`` (fold -> normalise -> unfold -> GELU as two stencil
                 kernels on tap-major columns), GEMM
"""
import torch
import torch.nn.functional as F

from ... import config, ops
from ...nn_util import as_nchw, as_pm, cl, conv
from ...window_index import padded_grid, token_grid, window_key_table

WIN = (5, 9)
POOL = (4, 4)
KS, ST, PD = 7, 3, 3


class TransformerExec:
    def __init__(self, net, depths=8, hidden=512, channel=128, ffn_ch=40):
        self.net, self.depths, self.hidden, self.channel, self.ffn_ch = net, depths, hidden, channel, ffn_ch

    # ------------------------------------------------------------------ packed weights
    def _ss(self):
        def build():
            P = self.net.P
            return cl(P["ss.embedding.weight"].view(self.hidden, self.channel, KS, KS)), P["ss.embedding.bias"].contiguous()
        return self.net.packed("ss", build)

    def _sc(self):
        def build():
            P = self.net.P
            w = P["sc.embedding.weight"].view(self.channel, KS, KS, self.hidden).permute(3, 0, 1, 2).contiguous()
            return w
        return self.net.packed("sc", build)

    def _sc_bias_map(self, hw):
        """fold of the Linear bias: constant per (h,w); sparse_transformer.py:52-59."""
        def build():
            fh, fw = token_grid(hw)
            b = self.net.P["sc.embedding.bias"].view(1, -1, 1).expand(1, -1, fh * fw)
            return F.fold(b, hw, (KS, KS), stride=ST, padding=PD).contiguous(memory_format=torch.channels_last)
        return self.net.packed(f"scb:{hw}", build)

    def _qkv(self, i):
        def build():
            P, p = self.net.P, f"transformers.transformer.{i}.attention."
            w = torch.cat([P[p + "query.weight"], P[p + "key.weight"], P[p + "value.weight"]], 0).contiguous()
            b = torch.cat([P[p + "query.bias"], P[p + "key.bias"], P[p + "value.bias"]], 0).contiguous()
            wkv = torch.cat([P[p + "key.weight"], P[p + "value.weight"]], 0).contiguous()
            bkv = torch.cat([P[p + "key.bias"], P[p + "value.bias"]], 0).contiguous()
            wpool = P[p + "pool_layer.weight"]                                     # [C,1,kh,kw] -> tap-major [kh*kw, C]
            return w, b, wkv, bkv, wpool.reshape(wpool.shape[0], -1).t().contiguous(), P[p + "pool_layer.bias"].contiguous()
        return self.net.packed(f"qkv{i}", build)

    def _ffn(self, i):
        def build():
            P, p = self.net.P, f"transformers.transformer.{i}.mlp."
            ch = self.ffn_ch
            perm = torch.arange(49 * ch, device=P[p + "fc1.0.weight"].device).view(ch, 49).t().reshape(-1)
            return (P[p + "fc1.0.weight"][perm].contiguous(), P[p + "fc1.0.bias"][perm].contiguous(),
                    P[p + "fc2.1.weight"][:, perm].contiguous(), P[p + "fc2.1.bias"].contiguous())
        return self.net.packed(f"ffn{i}", build)

    # ------------------------------------------------------------------ soft split / composition
    def soft_split(self, feat):
        """feat [t,c,h,w] channels_last -> tokens [t,fh,fw,hidden] pixel-major (SoftSplit.forward :19-31)."""
        return as_pm(conv(feat, self._ss(), ST, PD))

    def soft_comp(self, tokens, hw, res=None):
        """tokens [t,fh,fw,hidden] -> [t,c,h,w] (SoftComp.forward :49-61); `res` = skip added in the last conv's epilogue."""
        fh, fw = tokens.shape[1:3]
        op = (hw[0] + 2 - 3 * fh, hw[1] + 2 - 3 * fw)
        y = F.conv_transpose2d(as_nchw(tokens), self._sc(), None, stride=ST, padding=PD, output_padding=op)
        y = y + self._sc_bias_map(tuple(hw))
        P = self.net.P
        return conv(y, self.net.packed("scbc", lambda: (cl(P["sc.bias_conv.weight"]), P["sc.bias_conv.bias"].contiguous())), 1, 1,
                    res=res)

    # ------------------------------------------------------------------ transformer
    def run(self, tokens, hw, flags, t_dilation=2):
        """tokens [t,fh,fw,C]; flags int32 [n_windows] (1 = masked window).  :294-344."""
        assert self.depths % t_dilation == 0, "wrong t_dilation input."
        with config.linear_precision():
            return self._run(tokens, hw, flags, t_dilation)

    def _run(self, tokens, hw, flags, t_dilation):
        t, fh, fw, C = tokens.shape
        H2, W2 = padded_grid(fh, fw, WIN)
        NT = H2 * W2
        key_tok = self.net.packed(f"ktab:{H2}x{W2}", lambda: torch.from_numpy(window_key_table(H2, W2, WIN)).to(tokens.device))
        P = self.net.P
        x = tokens.contiguous()
        pad = (H2 != fh) or (W2 != fw)
        norm = lambda i, k: (P[f"transformers.transformer.{i}.norm{k}.weight"], P[f"transformers.transformer.{i}.norm{k}.bias"])
        _, y = ops.add_layernorm(x, None, *norm(0, 1))
        for i in range(self.depths):
            p = f"transformers.transformer.{i}."
            wqkv, bqkv, wkv, bkv, wpool, bpool = self._qkv(i)
            if pad:                                                    # zeros are padded *before* q/k/v (:168-176)
                y = F.pad(y, (0, 0, 0, W2 - fw, 0, H2 - fh))
            qkv = F.linear(y, wqkv, bqkv).view(t, NT, 3 * C)
            pooled = ops.pool_depthwise(y, wpool, bpool, POOL[0], POOL[1])                      # [t,ph,pw,C]
            pool_kv = F.linear(pooled.reshape(t, -1, C), wkv, bkv)                              # [t,NP,2C]
            att = ops.sparse_window_attn(qkv, pool_kv, key_tok, flags, t, NT, i % t_dilation, t_dilation,
                                         WN=WIN[0] * WIN[1], C=C).view(t, H2, W2, C)
            if pad:
                att = att[:, :fh, :fw]
            # x = x + proj(att); y = norm2(x)   (residual add fused into the LayerNorm pass)
            x, y = ops.add_layernorm(x, F.linear(att, P[p + "attention.proj.weight"], P[p + "attention.proj.bias"]), *norm(i, 2))
            w1, b1, w2, b2 = self._ffn(i)
            hdn = F.linear(y.reshape(t * fh * fw, C), w1, b1)
            hdn = ops.ffn_overlap_add(hdn, t, hw[0], hw[1], self.ffn_ch)
            d = F.linear(hdn, w2, b2).view(t, fh, fw, C)
            if i + 1 < self.depths:                                    # x = x + mlp(y); y = norm1 of the next block
                x, y = ops.add_layernorm(x, d, *norm(i + 1, 1))
            else:
                x = x + d
        return x
