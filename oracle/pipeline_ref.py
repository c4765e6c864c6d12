"""Oracle (test infrastructure): the inference driver loop, restated as a function.

Follows inference_propainter.py:298-452 (the script is not importable: its logic sits
under ``__main__`` and it needs imageio) and get_ref_index :159-173.  Tensors in,
uint8 frames out; file I/O, resize and mask dilation (:26-156) are outside this path.
"""
import numpy as np
import torch

from . import flowcomp_ref, generator_ref, raft_ref


def get_ref_index(mid, neighbor_ids, length, ref_stride=10, ref_num=-1):
    """inference_propainter.py:159-173."""
    out = []
    if ref_num == -1:
        return [i for i in range(0, length, ref_stride) if i not in neighbor_ids]
    lo = max(0, mid - ref_stride * (ref_num // 2))
    hi = min(length, mid + ref_stride * (ref_num // 2))
    for i in range(lo, hi, ref_stride):
        if i not in neighbor_ids:
            if len(out) > ref_num:
                break
            out.append(i)
    return out


def raft_clip_len(width):
    """inference_propainter.py:302-309."""
    return 12 if width <= 640 else 8 if width <= 720 else 4 if width <= 1280 else 2


def to_float_frames(frames_u8):
    """core/utils.py:130-170 (ToTorchFormatTensor: HWC uint8 -> CHW float /255) followed by
    inference_propainter.py:264 (*2-1).  frames_u8 [T,H,W,3] uint8 -> [1,T,3,H,W] float32."""
    x = torch.as_tensor(frames_u8).permute(0, 3, 1, 2).contiguous().float().div(255)
    return (x * 2 - 1).unsqueeze(0)


def _ec(on, t):
    """the reference's torch.cuda.empty_cache() calls (inference_propainter.py:323,360,395,452); no-op on CPU tensors"""
    if on and t.is_cuda:
        torch.cuda.empty_cache()


def stage_flow(sds, frames, raft_iter=20, empty_cache=False):
    """:302-330."""
    T, W = frames.shape[1], frames.shape[-1]
    clip = raft_clip_len(W)
    if T <= clip:
        return raft_ref.raft_bi(sds["raft"], frames, raft_iter)
    ff, bb = [], []
    for f in range(0, T, clip):
        e = min(T, f + clip)
        a, b = raft_ref.raft_bi(sds["raft"], frames[:, max(f - 1, 0):e] if f else frames[:, f:e], raft_iter)
        ff.append(a)
        bb.append(b)
        _ec(empty_cache, frames)
    return torch.cat(ff, 1), torch.cat(bb, 1)


def stage_complete(sds, flows_bi, flow_masks, subvideo_length=80, empty_cache=False):
    """:341-368."""
    L = flows_bi[0].shape[1]
    sd = sds["rfc"]
    if L <= subvideo_length:
        pred = flowcomp_ref.forward_bidirect_flow(sd, flows_bi, flow_masks)
        return flowcomp_ref.combine_flow(flows_bi, pred, flow_masks)
    pf, pb, pad = [], [], 5
    for f in range(0, L, subvideo_length):
        s, e = max(0, f - pad), min(L, f + subvideo_length + pad)
        ps, pe = f - s, e - min(L, f + subvideo_length)
        sub = (flows_bi[0][:, s:e], flows_bi[1][:, s:e])
        pred = flowcomp_ref.forward_bidirect_flow(sd, sub, flow_masks[:, s:e + 1])
        pred = flowcomp_ref.combine_flow(sub, pred, flow_masks[:, s:e + 1])
        pf.append(pred[0][:, ps:e - s - pe])
        pb.append(pred[1][:, ps:e - s - pe])
        _ec(empty_cache, flow_masks)
    return torch.cat(pf, 1), torch.cat(pb, 1)


def stage_img_prop(frames, masks_dilated, pred_flows, subvideo_length=80, empty_cache=False):
    """:371-404."""
    T = frames.shape[1]
    masked = frames * (1 - masks_dilated)
    sub = min(100, subvideo_length)
    if T <= sub:
        prop, um = generator_ref.img_propagation(masked, pred_flows[0], pred_flows[1], masks_dilated, "nearest")
        return frames * (1 - masks_dilated) + prop * masks_dilated, um
    uf, umk, pad = [], [], 10
    for f in range(0, T, sub):
        s, e = max(0, f - pad), min(T, f + sub + pad)
        ps, pe = f - s, e - min(T, f + sub)
        prop, um = generator_ref.img_propagation(masked[:, s:e], pred_flows[0][:, s:e - 1], pred_flows[1][:, s:e - 1],
                                                 masks_dilated[:, s:e], "nearest")
        upd = frames[:, s:e] * (1 - masks_dilated[:, s:e]) + prop * masks_dilated[:, s:e]
        uf.append(upd[:, ps:e - s - pe])
        umk.append(um[:, ps:e - s - pe])
        _ec(empty_cache, frames)
    return torch.cat(uf, 1), torch.cat(umk, 1)


def window_plan(T, neighbor_length=10, ref_stride=10, subvideo_length=80):
    """:406-421: the (neighbor_ids, ref_ids) list of the sliding-window loop."""
    ns = neighbor_length // 2
    ref_num = subvideo_length // ref_stride if T > subvideo_length else -1
    plan = []
    for f in range(0, T, ns):
        nb = list(range(max(0, f - ns), min(T, f + ns + 1)))
        plan.append((nb, get_ref_index(f, nb, T, ref_stride, ref_num)))
    return plan


def stage_generate(sds, upd_frames, masks_dilated, upd_masks, pred_flows, ori_u8, neighbor_length=10,
                   ref_stride=10, subvideo_length=80, empty_cache=False):
    """:406-452 incl. uint8 truncation, masked composite and the order-dependent 1/2-1/2 blend."""
    T = upd_frames.shape[1]
    comp = [None] * T
    for nb, refs in window_plan(T, neighbor_length, ref_stride, subvideo_length):
        ids = nb + refs
        fl = (pred_flows[0][:, nb[:-1]], pred_flows[1][:, nb[:-1]])
        pred = generator_ref.generator_forward(sds["gen"], upd_frames[:, ids], fl, masks_dilated[:, ids],
                                               upd_masks[:, ids], len(nb))
        pred = ((pred[0] + 1) / 2).cpu().permute(0, 2, 3, 1).numpy() * 255
        bm = masks_dilated[0, nb].cpu().permute(0, 2, 3, 1).numpy().astype(np.uint8)
        for i, idx in enumerate(nb):
            img = np.array(pred[i]).astype(np.uint8) * bm[i] + ori_u8[idx] * (1 - bm[i])
            if comp[idx] is None:
                comp[idx] = img
            else:
                comp[idx] = comp[idx].astype(np.float32) * 0.5 + img.astype(np.float32) * 0.5
            comp[idx] = comp[idx].astype(np.uint8)
        _ec(empty_cache, upd_frames)
    return np.stack(comp, 0)


def run_pipeline(sds, frames_u8, flow_masks, masks_dilated, raft_iter=20, neighbor_length=10, ref_stride=10,
                 subvideo_length=80, return_stages=False, empty_cache=False):
    """Whole path.  frames_u8 [T,H,W,3] uint8 (numpy); masks [1,T,1,H,W] float {0,1}."""
    frames = to_float_frames(frames_u8).to(masks_dilated.device)
    with torch.no_grad():
        gt = stage_flow(sds, frames, raft_iter, empty_cache)
        _ec(empty_cache, frames)                                           # :330
        pred = stage_complete(sds, gt, flow_masks, subvideo_length, empty_cache)
        _ec(empty_cache, frames)                                           # :368
        upd_f, upd_m = stage_img_prop(frames, masks_dilated, pred, subvideo_length, empty_cache)
        _ec(empty_cache, frames)                                           # :404
        comp = stage_generate(sds, upd_f, masks_dilated, upd_m, pred, np.asarray(frames_u8), neighbor_length,
                              ref_stride, subvideo_length, empty_cache)
    if return_stages:
        return comp, {"gt_flows": gt, "pred_flows": pred, "updated_frames": upd_f, "updated_masks": upd_m}
    return comp
