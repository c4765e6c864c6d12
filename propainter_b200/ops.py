"""Tensor-level wrappers over the C ABI.  Each takes/returns torch CUDA tensors, passes raw pointers +
the current stream, and raises on any non-zero return code.  fp32 only."""
import ctypes
import math

import torch

from . import _lib
from ._lib import PPAttnParams, PPConvParams, PPWindowIds, check

LOG2E = 1.4426950408889634

# number of kernels of libpropainter_b200.so launched so far (bench.py reports the delta as gpu_launches)
LAUNCHES = 0


def _count(n):
    global LAUNCHES
    LAUNCHES += n


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t, dtype=torch.float32):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("propainter_b200 ops need CUDA tensors (there is no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"expected {dtype}, got {t.dtype}")
    return ctypes.c_void_p(t.data_ptr())


def _dense(t):
    if not t.is_contiguous():
        raise RuntimeError("tensor must be contiguous")
    return t


def _pm(t):
    """pixel-major view [..., C] whose pixels are `ld` elements apart -> (ptr, ld)."""
    if t.stride(-1) != 1:
        raise RuntimeError("channel dim must be unit-stride")
    ld = t.stride(-2)
    exp = ld
    for d in range(t.dim() - 2, -1, -1):
        if t.shape[d] != 1 and t.stride(d) != exp:
            raise RuntimeError("pixel-major view must be dense over its pixels")
        exp *= t.shape[d]
    return _p(t), ld


def corr_ld(w):
    return (w + 3) & ~3


def corr_alloc(n_pairs, h, w, device):
    """Pyramid buffers [n_pairs*h*w, h>>l, ld_l] (zero-filled once so the row padding is defined)."""
    lv, hl, wl = [], h, w
    for _ in range(4):
        lv.append(torch.zeros(n_pairs * h * w, hl, corr_ld(wl), device=device, dtype=torch.float32))
        hl, wl = hl // 2, wl // 2
    return lv


def _level_array(levels):
    return (ctypes.c_void_p * 4)(*[lv.data_ptr() for lv in levels])


def corr_build(fmap, idx1, idx2, levels, h, w):
    """fmap [frames, h*w, D] pixel-major; idx1/idx2 int32 CUDA [n_pairs]; fills the 4 pyramid levels."""
    L = _lib.lib()
    n_pairs = idx1.numel()
    D = fmap.shape[-1]
    check(L.pp_corr_build(_p(_dense(fmap)), D, _p(idx1, torch.int32), _p(idx2, torch.int32), n_pairs, _p(levels[0]), h, w,
                          _stream()), "pp_corr_build")
    check(L.pp_corr_pool_pyramid(_level_array(levels), n_pairs * h * w, h, w, _stream()), "pp_corr_pool_pyramid")
    _count(4)


def corr_lookup(levels, coords, out=None, tma=True):
    """coords [B,h,w,2] -> [B,h,w,324].  tma=False selects the plain-load baseline kernel."""
    B, h, w, _ = coords.shape
    if out is None:
        out = torch.empty(B, h, w, 324, device=coords.device, dtype=torch.float32)
    fn = _lib.lib().pp_corr_lookup if tma else _lib.lib().pp_corr_lookup_ldg
    check(fn(_level_array(levels), _p(_dense(coords)), _p(_dense(out)), B, h, w, _stream()), "pp_corr_lookup")
    _count(1)
    return out


def convex_upsample(mask_pm, flow_lr, mask_scale=0.25):
    """mask_pm [n,h,w,576] pixel-major, flow_lr [n,h,w,2] -> planar [n,2,8h,8w]."""
    n, h, w, _ = flow_lr.shape
    mp, ld = _pm(mask_pm)
    out = torch.empty(n, 2, 8 * h, 8 * w, device=flow_lr.device, dtype=torch.float32)
    check(_lib.lib().pp_convex_upsample(mp, ld, mask_scale, _p(_dense(flow_lr)), _p(out), n, h, w, _stream()),
          "pp_convex_upsample")
    _count(1)
    return out


def img_prop_scan(frames, flows_f, flows_b, masks, nearest=True):
    """frames [t,3,H,W], flows [t-1,2,H,W], masks [t,1,H,W] (planar) -> (frames_out, masks_out)."""
    L = _lib.lib()
    t, _, H, W = frames.shape
    ws_bytes = L.pp_img_prop_scan_workspace_bytes(t, H, W)
    ws = torch.empty(ws_bytes // 4, device=frames.device, dtype=torch.float32)
    of, om = torch.empty_like(frames), torch.empty_like(masks)
    check(L.pp_img_prop_scan(_p(_dense(frames)), _p(_dense(flows_f)), _p(_dense(flows_b)), _p(_dense(masks)), _p(of),
                             _p(om), _p(ws), ws_bytes, t, H, W, int(bool(nearest)), _stream()), "pp_img_prop_scan")
    _count(2 * (t - 1))
    return of, om


def prop_cond(cur, prop, fprop, fcheck, mcur, cond, bb, first):
    """cur/prop [h,w,C] pixel-major views; fprop/fcheck/mcur [h,w,2]; cond [h,w,ldc]; bb [h,w,ldb]."""
    h, w, C = cur.shape
    cp, ldc = _pm(cur)
    pp_, ldp = _pm(prop) if prop is not None else (None, ldc)
    cdp, ldcd = _pm(cond) if cond is not None else (None, 2 * C + 8)
    bp, ldb = _pm(bb)
    check(_lib.lib().pp_prop_cond(cp, ldc, pp_, ldp, _p(fprop), _p(fcheck), _p(_dense(mcur)), cdp, ldcd, bp, ldb, h, w,
                                  C, int(bool(first)), _stream()), "pp_prop_cond")
    _count(1)


def deform_align(x, o, flow, max_res, w_packed, bias, out, o_bias=None):
    """x [H,W,Cin] view, o [H,W,>=432] view (raw conv_offset.6 output; its bias may be passed as o_bias instead of being
    pre-added), flow [H,W,2]|None, out [H,W,128] view (all pixel-major).  The warp-level mma.sync implementation: kept as
    the measured baseline of deform_gather + conv_umma (config.UMMA_CONV)."""
    H, W, Cin = x.shape
    xp, ldx = _pm(x)
    op, ldo = _pm(o)
    outp, ldout = _pm(out)
    L = _lib.lib()
    ws_bytes = L.pp_deform_align_workspace_bytes(H, W)
    ws = torch.empty(max(ws_bytes // 4, 4), device=x.device, dtype=torch.float32)
    check(L.pp_deform_align(xp, ldx, op, ldo, _p(o_bias), _p(flow), float(max_res), _p(_dense(w_packed)), _p(bias), outp, ldout,
                            H, W, Cin, out.shape[-1], _p(ws), ws_bytes, _stream()), "pp_deform_align")
    _count(3)                                                    # tap pre-pass, GEMM, (split-K reduce)
    return out


def tf32_round(w):
    """fp32 -> nearest TF32 value (ties away from zero, = cvt.rna.tf32.f32), still stored as fp32.  tcgen05 kind::tf32
    ignores the low 13 mantissa bits; rounding once at pack time keeps the products unbiased."""
    i = w.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def pack_conv_weight(weight, seg_channels=None):
    """[Cout,Cin,KH,KW] -> [Cout,K] for pp_conv2d_umma: the input channels are split into the segments `seg_channels`
    (default: one segment), every segment is zero-padded to 32-channel blocks and K runs ((blk*KH + dy)*KW + dx)*32 + c."""
    Cout, Cin, KH, KW = weight.shape
    seg_channels = [Cin] if seg_channels is None else list(seg_channels)
    if sum(seg_channels) != Cin:
        raise RuntimeError("pack_conv_weight: segments do not add up to Cin")
    blocks, c0 = [], 0
    for C in seg_channels:
        for b in range(0, C, 32):
            cw = min(32, C - b)
            blk = weight.new_zeros(Cout, KH, KW, 32)
            blk[..., :cw] = weight[:, c0 + b:c0 + b + cw].permute(0, 2, 3, 1)
            blocks.append(blk.reshape(Cout, KH * KW * 32))
        c0 += C
    return tf32_round(torch.cat(blocks, 1).contiguous())


def _pm4(t):
    """[n,H,W,C] pixel-major view (unit channel stride, dense over n*H*W pixels) -> (ptr, ld)."""
    if t.dim() != 4:
        raise RuntimeError("expected [n,H,W,C]")
    return _pm(t)


def conv_umma(segs, w_packed, KH, KW, Cout, bias=None, act="none", slope=0.0, pre=None, res=None, post_relu=False, out=None,
              round_tf32=False, bn=0, tile_w=0, tile_m=0):
    """tcgen05 implicit-GEMM conv (stride 1, same padding) with fused epilogue.  segs: list of [n,H,W,C_i] pixel-major
    views = the channel-concatenated input; w_packed from pack_conv_weight(weight, [C_i...]); pre / res / out
    [n,H,W,Cout] views (channel slices of wider buffers allowed).  Returns out."""
    n, H, W, _ = segs[0].shape
    if out is None:
        out = torch.empty(n, H, W, Cout, device=segs[0].device, dtype=torch.float32)
    prm = PPConvParams()
    prm.nseg = len(segs)
    kblocks = 0
    for i, sgm in enumerate(segs):
        if tuple(sgm.shape[:3]) != (n, H, W):
            raise RuntimeError("conv_umma: segment shape mismatch")
        ptr, ld = _pm4(sgm)
        prm.seg[i].x, prm.seg[i].ld, prm.seg[i].C = ptr.value, ld, sgm.shape[-1]
        kblocks += (sgm.shape[-1] + 31) // 32
    if tuple(w_packed.shape) != (Cout, kblocks * KH * KW * 32):
        raise RuntimeError(f"conv_umma: packed weight {tuple(w_packed.shape)} does not match {(Cout, kblocks * KH * KW * 32)}")
    prm.n, prm.H, prm.W, prm.KH, prm.KW = n, H, W, KH, KW
    prm.w_packed, prm.Cout = _p(_dense(w_packed)).value, Cout
    prm.bias = _p(bias).value if bias is not None else None
    for name, t in (("pre", pre), ("res", res), ("out", out)):
        if t is None:
            setattr(prm, name, None)
            setattr(prm, "ld_" + name, 0)
            continue
        if tuple(t.shape) != (n, H, W, Cout):
            raise RuntimeError(f"conv_umma: {name} shape {tuple(t.shape)} != {(n, H, W, Cout)}")
        ptr, ld = _pm4(t)
        setattr(prm, name, ptr.value)
        setattr(prm, "ld_" + name, ld)
    prm.act, prm.slope, prm.post_relu, prm.round_tf32 = ACT[act], float(slope), int(bool(post_relu)), int(bool(round_tf32))
    prm.bn, prm.tile_w, prm.tile_m = int(bn), int(tile_w), int(tile_m)
    check(_lib.lib().pp_conv2d_umma(ctypes.byref(prm), _stream()), "pp_conv2d_umma")
    _count(1)
    return out


def deform_gather(x, o, flow, max_res, cols=None, o_bias=None, x2=None):
    """x [n,H,W,Cin] view (or, with x2, the two halves x | x2 of Cin/2 channels each), o [n,H,W,>=432] raw conv_offset
    output, flow [n,H,W,2] | None -> cols [n,H,W,9*Cin] (modulated bilinear samples, k*Cin + c, TF32-rounded): the A
    operand of the deformable conv's GEMM."""
    n, H, W, Cin = x.shape
    xp, ldx = _pm4(x)
    x2p, ldx2 = (None, 0)
    if x2 is not None:
        if x2.shape != x.shape:
            raise RuntimeError("deform_gather: x2 must have the shape of x")
        x2p, ldx2 = _pm4(x2)
        Cin *= 2
    op, ldo = _pm4(o)
    if cols is None:
        cols = torch.empty(n, H, W, 9 * Cin, device=x.device, dtype=torch.float32)
    check(_lib.lib().pp_deform_gather(xp, ldx, x2p, ldx2, op, ldo, _p(o_bias), _p(_dense(flow)) if flow is not None else None,
                                      float(max_res), _p(_dense(cols)), n, H, W, Cin, _stream()), "pp_deform_gather")
    _count(1)
    return cols


def flow_warp_fbcheck(feat, fprop, fcheck=None, warped=None, aux=None, want_warp=True, round_tf32=False):
    """flow_warp (bilinear) of pixel-major maps feat [n,h,w,C] by fprop [n,h,w,2] -> warped [n,h,w,C] (views allowed);
    with fcheck also the forward-backward validity: aux [n,h,w,>=3] view receives (fx, fy, valid).  Returns (warped, aux)."""
    n, h, w = fprop.shape[:3]
    fp_, ldf, wp_, ldw, C = None, 0, None, 0, 0
    if want_warp:
        C = feat.shape[-1]
        if warped is None:
            warped = torch.empty(n, h, w, C, device=feat.device, dtype=torch.float32)
        fp_, ldf = _pm4(feat)
        wp_, ldw = _pm4(warped)
    ap, lda = (None, 0)
    if fcheck is not None:
        if aux is None:
            aux = torch.empty(n, h, w, 4, device=fprop.device, dtype=torch.float32)
        ap, lda = _pm4(aux)
    check(_lib.lib().pp_flow_warp_fbcheck(fp_, ldf, _p(_dense(fprop)), _p(_dense(fcheck)) if fcheck is not None else None, wp_, ldw,
                                          ap, lda, n, h, w, C, int(bool(round_tf32)), _stream()), "pp_flow_warp_fbcheck")
    _count(1)
    return warped, aux


def pack_deform_weight_umma(weight):
    """deform-conv weight [Cout,Cin,3,3] -> [Cout, 9*Cin] with k = tap*Cin + c (the column order of deform_gather), TF32."""
    co, ci = weight.shape[:2]
    return tf32_round(weight.permute(0, 2, 3, 1).reshape(co, 9 * ci).contiguous())


def pack_deform_weight(weight):
    """[Cout,Cin,3,3] -> [9*Cin, Cout], row = tap*Cin + c."""
    co, ci = weight.shape[:2]
    return weight.permute(2, 3, 1, 0).reshape(9 * ci, co).contiguous()


def gen_prep(flows_f, flows_b, masks_in, masks_upd, lt):
    """planar flows [lt-1,2,H,W], masks [>=lt,1,H,W] -> dsf, dsb [lt-1,h,w,2], pmask [lt,h,w,2]."""
    H, W = masks_in.shape[-2:]
    h, w = H // 4, W // 4
    dev = masks_in.device
    dsf = torch.empty(max(lt - 1, 1), h, w, 2, device=dev, dtype=torch.float32)
    dsb = torch.empty_like(dsf)
    pmask = torch.empty(lt, h, w, 2, device=dev, dtype=torch.float32)
    check(_lib.lib().pp_gen_prep(_p(_dense(flows_f)), _p(_dense(flows_b)), _p(_dense(masks_in)), _p(_dense(masks_upd)),
                                 _p(dsf), _p(dsb), _p(pmask), lt, H, W, _stream()), "pp_gen_prep")
    _count(1)
    return dsf[:lt - 1], dsb[:lt - 1], pmask


def window_mask(pmask, fh, fw, nwh, nww):
    lt, h, w, _ = pmask.shape
    flags = torch.empty(nwh * nww, device=pmask.device, dtype=torch.int32)
    check(_lib.lib().pp_window_mask(_p(_dense(pmask)), lt, h, w, fh, fw, nwh, nww, _p(flags, torch.int32), _stream()),
          "pp_window_mask")
    _count(1)
    return flags


def sparse_window_attn(qkv, pool_kv, key_tok, flags, t, NT, kf_start, kf_step, out=None, WN=45, C=512, impl="umma"):
    """qkv [t,NT,3C]; pool_kv [t,NP,2C]; key_tok int32 [nwin,NKO]; flags int32 [nwin] -> out [t,NT,C].
    impl: "umma" = tcgen05/TMEM kernel for masked windows (default), "mma" = warp-level mma.sync baseline."""
    if out is None:
        out = torch.empty(t, NT, C, device=qkv.device, dtype=torch.float32)
    prm = PPAttnParams()
    prm.qkv, prm.pool = qkv.data_ptr(), pool_kv.data_ptr()
    prm.key_tok, prm.flags, prm.out = key_tok.data_ptr(), flags.data_ptr(), out.data_ptr()
    prm.ld_qkv, prm.ld_pool, prm.ld_out = qkv.stride(-2), pool_kv.stride(-2), out.stride(-2)
    prm.t, prm.NT, prm.WN, prm.NKO, prm.NP, prm.C = t, NT, WN, key_tok.shape[1], pool_kv.shape[1], C
    prm.kf_start, prm.kf_step = kf_start, kf_step
    prm.nkf = len(range(kf_start, t, kf_step))
    if prm.nkf == 0:
        out.zero_()                 # empty key set: masked windows yield zeros (softmax over an empty dim), see the C entry
    prm.scale_log2 = LOG2E / math.sqrt(128.0)
    for tns, dt in ((qkv, torch.float32), (pool_kv, torch.float32), (key_tok, torch.int32), (flags, torch.int32)):
        _p(_dense(tns), dt)
    fn = _lib.lib().pp_sparse_window_attn if impl == "umma" else _lib.lib().pp_sparse_window_attn_mma
    check(fn(ctypes.byref(prm), key_tok.shape[0], _stream()), "pp_sparse_window_attn")
    _count(2)
    return out


def ffn_overlap_add(Y, frames, h, w, CH=40):
    """Y [frames*fh*fw, 49*CH] (tap-major columns) -> gelu(unfold(fold(Y)/norm)) same shape."""
    L = _lib.lib()
    Z = torch.empty_like(Y)
    ws_bytes = L.pp_ffn_overlap_add_workspace_bytes(frames, h, w, CH)
    ws = torch.empty(ws_bytes // 4, device=Y.device, dtype=torch.float32)
    check(L.pp_ffn_overlap_add(_p(_dense(Y)), Y.shape[-1], _p(Z), Z.shape[-1], frames, h, w, CH, _p(ws), ws_bytes,
                               _stream()), "pp_ffn_overlap_add")
    _count(2)
    return Z


def gru_gate(zr_pm, bias, net_view, z_out, rnet_view, pre=None):
    """zr_pm [..,2C] raw gate conv output; net_view / rnet_view: C-channel slices of HX / RX; z_out dense [..,C];
    bias [2C] / pre [..,2C] optional addends."""
    C = z_out.shape[-1]
    np_, ldn = _pm(net_view)
    rp, ldr = _pm(rnet_view)
    check(_lib.lib().pp_gru_gate(_p(_dense(zr_pm)), _p(bias), _p(_dense(pre)) if pre is not None else None, np_, ldn,
                                 _p(_dense(z_out)), rp, ldr, z_out.numel() // C, C, _stream()), "pp_gru_gate")
    _count(1)


def gru_update(q_pm, bias, z, net_view, net_copy=None, pre=None):
    """h = (1-z)*h + z*tanh(q+bias+pre) in place on the state slice; `net_copy` (dense) also receives h."""
    C = z.shape[-1]
    np_, ldn = _pm(net_view)
    check(_lib.lib().pp_gru_update(_p(_dense(q_pm)), _p(bias), _p(_dense(pre)) if pre is not None else None, _p(_dense(z)), np_, ldn,
                                   _p(_dense(net_copy)) if net_copy is not None else None, z.numel() // C, C, _stream()),
          "pp_gru_update")
    _count(1)


def raft_pack_motion(mot_pm, flow_pm, d0_view, d1_view, bias=None):
    """mot_pm [..,128] (channels 126,127 ignored), flow_pm [..,2] -> 128-channel slot views of HX and RX.
    With `bias`, mot_pm is the raw conv output and relu(mot + bias) is applied on the way."""
    mp, ldm = _pm(mot_pm)
    p0, ld0 = _pm(d0_view)
    p1, ld1 = _pm(d1_view)
    if ld0 != ld1:
        raise RuntimeError("HX / RX must share the pixel stride")
    check(_lib.lib().pp_raft_pack_motion(mp, ldm, _p(bias), _p(_dense(flow_pm)), p0, p1, ld0, flow_pm.numel() // 2, _stream()),
          "pp_raft_pack_motion")
    _count(1)


ACT = {"none": 0, "relu": 1, "leaky": 2, "sigmoid": 3, "tanh": 4}


def bias_act(x_pm, bias=None, act="none", slope=0.0, res=None, post_relu=False, out=None, pre=None):
    """out = post(act(x + bias + pre) + res) on pixel-major views [..., C] (unit channel stride, dense over pixels; x / pre /
    res / out may each be a channel slice of a wider buffer).  out=None -> in place on x_pm.  Returns out."""
    C = x_pm.shape[-1]
    out = x_pm if out is None else out
    if out.shape != x_pm.shape or (res is not None and res.shape != x_pm.shape) or (pre is not None and pre.shape != x_pm.shape):
        raise RuntimeError("bias_act: shape mismatch")
    xp, ldx = _pm(x_pm)
    op, ldo = _pm(out)
    rp, ldr = _pm(res) if res is not None else (None, C)
    if pre is not None:
        pp, ldp = _pm(pre)
        check(_lib.lib().pp_bias_act_pre(xp, ldx, _p(bias), pp, ldp, rp, ldr, op, ldo, x_pm.numel() // C, C, ACT[act], float(slope),
                                         int(bool(post_relu)), _stream()), "pp_bias_act_pre")
    else:
        check(_lib.lib().pp_bias_act(xp, ldx, _p(bias), rp, ldr, op, ldo, x_pm.numel() // C, C, ACT[act], float(slope),
                                     int(bool(post_relu)), _stream()), "pp_bias_act")
    _count(1)
    return out


def bias_act_(x_pm, bias, act="none", slope=0.0):
    """in-place act(x + bias); returns x_pm."""
    return bias_act(x_pm, bias, act, slope)


def pool_depthwise(x_pm, w_taps, bias, kh, kw):
    """depthwise conv, kernel = stride = (kh,kw): x_pm [n,H,W,C] -> [n,H//kh,W//kw,C]; w_taps [kh*kw, C]."""
    n, H, W, C = x_pm.shape
    xp, ld = _pm(x_pm)
    out = torch.empty(n, (H - kh) // kh + 1, (W - kw) // kw + 1, C, device=x_pm.device, dtype=torch.float32)
    check(_lib.lib().pp_pool_depthwise(xp, ld, _p(_dense(w_taps)), _p(bias), _p(out), n, H, W, C, kh, kw, _stream()),
          "pp_pool_depthwise")
    _count(1)
    return out


def add_layernorm(x, delta, gamma, beta, eps=1e-5):
    """(x + delta, LayerNorm(x + delta)) over the last dim; delta=None -> (x, LayerNorm(x)).  Dense tensors."""
    C = x.shape[-1]
    y = torch.empty_like(x)
    xo = torch.empty_like(x) if delta is not None else None
    check(_lib.lib().pp_add_layernorm(_p(_dense(x)), _p(_dense(delta)) if delta is not None else None, _p(gamma), _p(beta),
                                      _p(xo), _p(y), x.numel() // C, C, float(eps), _stream()), "pp_add_layernorm")
    _count(1)
    return (xo if delta is not None else x), y


def instance_norm(x_pm, relu=False, res=None, post_relu=False, eps=1e-5, out=None):
    """InstanceNorm2d(affine=False) on a dense pixel-major map [n,h,w,C] (+ ReLU, + residual add, + final ReLU)."""
    n, h, w, C = x_pm.shape
    out = torch.empty_like(x_pm) if out is None else out
    lib = _lib.lib()
    nbytes = lib.pp_instance_norm_workspace_bytes(n, h * w, C)
    ws = torch.empty(max(nbytes // 4, 4), device=x_pm.device, dtype=torch.float32)
    check(lib.pp_instance_norm(_p(_dense(x_pm)), _p(_dense(res)) if res is not None else None, _p(_dense(out)), n, h * w, C,
                               float(eps), int(bool(relu)), int(bool(post_relu)), _p(ws), ws.numel() * 4, _stream()),
          "pp_instance_norm")
    _count(2)
    return out


def upsample2x(x_pm):
    """pixel-major [n,h,w,C] -> [n,2h,2w,C], bilinear, align_corners=True."""
    n, h, w, C = x_pm.shape
    out = torch.empty(n, 2 * h, 2 * w, C, device=x_pm.device, dtype=torch.float32)
    check(_lib.lib().pp_upsample2x_bilinear(_p(_dense(x_pm)), _p(out), n, h, w, C, _stream()), "pp_upsample2x_bilinear")
    _count(1)
    return out


def mask_dilate(masks_u8, iterations):
    """uint8 [T,H,W] (non-zero = hole) -> float {0,1} [T,1,H,W], dilated `iterations` times with the 3x3 cross."""
    T, H, W = masks_u8.shape
    out = torch.empty(T, 1, H, W, device=masks_u8.device, dtype=torch.float32)
    check(_lib.lib().pp_mask_dilate(_p(_dense(masks_u8), torch.uint8), _p(out), T, H, W, int(iterations), _stream()),
          "pp_mask_dilate")
    _count(1)
    return out


def u8_to_frames(frames_u8):
    """uint8 [T,H,W,3] -> float planar [T,3,H,W] in [-1,1]."""
    T, H, W, _ = frames_u8.shape
    out = torch.empty(T, 3, H, W, device=frames_u8.device, dtype=torch.float32)
    check(_lib.lib().pp_u8_to_frames(_p(_dense(frames_u8), torch.uint8), _p(out), T, H, W, _stream()), "pp_u8_to_frames")
    _count(1)
    return out


def composite_blend(pred, masks, ori_u8, comp_u8, frame_ids, first_flags):
    """pred [n,3,H,W]; masks [T,1,H,W]; ori/comp uint8 [T,H,W,3] (comp updated in place)."""
    n, _, H, W = pred.shape
    ids = PPWindowIds()
    ids.n = n
    for i, (f, fl) in enumerate(zip(frame_ids, first_flags)):
        ids.frame[i], ids.first[i] = int(f), int(fl)
    check(_lib.lib().pp_composite_blend_u8(_p(_dense(pred)), _p(_dense(masks)), _p(_dense(ori_u8), torch.uint8),
                                           _p(_dense(comp_u8), torch.uint8), ctypes.byref(ids), H, W, _stream()),
          "pp_composite_blend_u8")
    _count(1)


# ---------------------------------------------------------------- resizing around the path (tables on the host, passes on the device)
_TABLES = {}


def resample_tables(kind, in_size, out_size, device=None, horizontal=True):
    """Per-axis tables of the library resamplers the reference calls (see include/propainter_b200.h), as torch tensors
    (on `device` if given).  kind: "bicubic" -> (bounds int32 [out,2], kk int32 [out,ksize]); "nearest" -> idx int32 [out];
    "linear_cv" -> (ofs int32 [out], coef int16 [out,2])."""
    key = (kind, in_size, out_size, str(device), horizontal)
    if key in _TABLES:
        return _TABLES[key]
    L = _lib.lib()
    if kind == "bicubic":
        ks = L.pp_resample_coeffs_bicubic(in_size, out_size, None, None, 0)
        bounds, kk = torch.empty(out_size, 2, dtype=torch.int32), torch.empty(out_size, ks, dtype=torch.int32)
        check(min(0, L.pp_resample_coeffs_bicubic(in_size, out_size, bounds.data_ptr(), kk.data_ptr(), kk.numel())), "pp_resample_coeffs_bicubic")
        res = (bounds, kk)
    elif kind == "nearest":
        idx = torch.empty(out_size, dtype=torch.int32)
        check(L.pp_resample_index_nearest(in_size, out_size, idx.data_ptr()), "pp_resample_index_nearest")
        res = (idx,)
    elif kind == "linear_cv":
        ofs, coef = torch.empty(out_size, dtype=torch.int32), torch.empty(out_size, 2, dtype=torch.int16)
        check(L.pp_resample_coeffs_linear_cv(in_size, out_size, int(bool(horizontal)), ofs.data_ptr(), coef.data_ptr()), "pp_resample_coeffs_linear_cv")
        res = (ofs, coef)
    else:
        raise ValueError(kind)
    if device is not None:
        res = tuple(t.to(device) for t in res)
    _TABLES[key] = res
    return res


def resize_frames_u8(frames_u8, size):
    """resize_frames (inference_propainter.py:34-45): uint8 [T,H,W,3] -> [T,Ho,Wo,3], size = (Wo, Ho) like PIL; = PIL.Image.resize(size)."""
    T, H, W, _ = frames_u8.shape
    Wo, Ho = size
    dev = frames_u8.device
    bx, kx = resample_tables("bicubic", W, Wo, dev)
    by, ky = resample_tables("bicubic", H, Ho, dev)
    out = torch.empty(T, Ho, Wo, 3, dtype=torch.uint8, device=dev)
    L = _lib.lib()
    nws = L.pp_resize_u8_bicubic_workspace_bytes(T, H, Wo)
    ws = torch.empty(max(nws, 16), dtype=torch.uint8, device=dev)
    check(L.pp_resize_u8_bicubic(_p(_dense(frames_u8), torch.uint8), _p(out, torch.uint8), T, H, W, Ho, Wo, _p(bx, torch.int32), _p(kx, torch.int32),
                                 kx.shape[1], _p(by, torch.int32), _p(ky, torch.int32), ky.shape[1], _p(ws, torch.uint8), ws.numel(), _stream()),
          "pp_resize_u8_bicubic")
    _count(2)
    return out


def resize_masks_u8(masks_u8, size):
    """mask_img.resize(size, Image.NEAREST) (inference_propainter.py:95-96): uint8 [T,H,W] -> [T,Ho,Wo]."""
    T, H, W = masks_u8.shape
    Wo, Ho = size
    dev = masks_u8.device
    (ix,), (iy,) = resample_tables("nearest", W, Wo, dev), resample_tables("nearest", H, Ho, dev)
    out = torch.empty(T, Ho, Wo, dtype=torch.uint8, device=dev)
    check(_lib.lib().pp_resize_u8_nearest(_p(_dense(masks_u8), torch.uint8), _p(out, torch.uint8), T, H, W, Ho, Wo, 1, _p(ix, torch.int32),
                                          _p(iy, torch.int32), _stream()), "pp_resize_u8_nearest")
    _count(1)
    return out


def resize_output_u8(frames_u8, size):
    """cv2.resize(f, out_size) of the composited frames (inference_propainter.py:469-470): uint8 [T,H,W,3] -> [T,Ho,Wo,3]."""
    T, H, W, _ = frames_u8.shape
    Wo, Ho = size
    dev = frames_u8.device
    xo, xa = resample_tables("linear_cv", W, Wo, dev, True)
    yo, ya = resample_tables("linear_cv", H, Ho, dev, False)
    out = torch.empty(T, Ho, Wo, 3, dtype=torch.uint8, device=dev)
    check(_lib.lib().pp_resize_u8_bilinear_cv(_p(_dense(frames_u8), torch.uint8), _p(out, torch.uint8), T, H, W, Ho, Wo, _p(xo, torch.int32),
                                              _p(xa, torch.int16), _p(yo, torch.int32), _p(ya, torch.int16), _stream()), "pp_resize_u8_bilinear_cv")
    _count(1)
    return out
