"""The inference driver: stage scheduling of inference_propainter.py:298-452 as a library call.

The reference keeps this logic under ``__main__`` with file I/O around it; here it is a class that
takes frames / masks as tensors (uint8 frames may come from pinned host memory), keeps everything
on the device, and replaces the per-window ``.cpu()`` + numpy compositing (:437-450) with one kernel,
so a clip needs exactly one host->device and one device->host copy.  Chunk boundaries, halo lengths,
reference-frame selection and the order-dependent blend follow the reference exactly.
"""
from dataclasses import dataclass

import torch

from . import ops
from .model.modules.flow_comp_raft import RAFT_bi
from .model.propainter import InpaintGenerator
from .model.recurrent_flow_completion import RecurrentFlowCompleteNet


@dataclass
class InferenceConfig:
    """argparse flags of inference_propainter.py:181-217 that shape the hot path."""
    raft_iter: int = 20
    ref_stride: int = 10
    neighbor_length: int = 10
    subvideo_length: int = 80
    # frames per RAFT call.  None = as many as an 8 GB correlation pyramid allows (never fewer than the reference's
    # 12/8/4/2, inference_propainter.py:302-309).  Frame pairs are independent, so this only changes batching.
    raft_clip_frames: int = None
    windows_in_flight: int = 3      # generator windows computed concurrently on separate streams (compositing stays ordered)
    # --fp16 (inference_propainter.py:211, :268-270, :323-330) halves the two nets and every tensor after RAFT.  The
    # kernels here compute in fp32 whatever the storage dtype (a `.half()` net is widened once, fp16 inputs are widened at
    # entry and results handed back in the caller's dtype), so the flag is accepted and changes nothing in the pipeline.
    fp16: bool = False


def get_ref_index(mid_neighbor_id, neighbor_ids, length, ref_stride=10, ref_num=-1):
    """inference_propainter.py:159-173 (incl. its `> ref_num` early-exit quirk)."""
    nb = set(neighbor_ids)
    if ref_num == -1:
        return [i for i in range(0, length, ref_stride) if i not in nb]
    half = ref_stride * (ref_num // 2)
    picked = []
    for i in range(max(0, mid_neighbor_id - half), min(length, mid_neighbor_id + half), ref_stride):
        if i in nb:
            continue
        if len(picked) > ref_num:
            break
        picked.append(i)
    return picked


def raft_clip_len(width):
    """:302-309"""
    for limit, n in ((640, 12), (720, 8), (1280, 4)):
        if width <= limit:
            return n
    return 2


def flow_chunks(T, clip):
    """Frame ranges [s,e) handed to RAFT_bi, with the 1-frame overlap of :314-319."""
    if T <= clip:
        return [(0, T)]
    return [(max(f - 1, 0), min(T, f + clip)) for f in range(0, T, clip)]


def halo_chunks(L, sub, pad):
    """(s, e, keep_lo, keep_hi) for the recompute-halo chunking of :342-364 / :373-398."""
    out = []
    for f in range(0, L, sub):
        s, e = max(0, f - pad), min(L, f + sub + pad)
        out.append((s, e, f - s, (e - s) - (e - min(L, f + sub))))
    return out


def window_plan(T, cfg):
    """[(neighbor_ids, ref_ids)] of the sliding-window loop :406-421."""
    stride = cfg.neighbor_length // 2
    ref_num = cfg.subvideo_length // cfg.ref_stride if T > cfg.subvideo_length else -1
    plan = []
    for f in range(0, T, stride):
        nb = list(range(max(0, f - stride), min(T, f + stride + 1)))
        plan.append((nb, get_ref_index(f, nb, T, cfg.ref_stride, ref_num)))
    return plan


def prepare_masks(masks_u8, mask_dilation=4, device="cuda"):
    """read_mask (inference_propainter.py:77-114) after file I/O: uint8 masks [T,H,W] (or [1,H,W], repeated by the
    caller) -> (flow_masks, masks_dilated), both float {0,1} [1,T,1,H,W] on the device.  Both use `mask_dilation`
    (the script passes args.mask_dilation for both, :238-240)."""
    m = masks_u8.to(device)
    d = ops.mask_dilate(m, mask_dilation).unsqueeze(0)
    return d, d.clone()


class ProPainterPipeline:
    """RAFT flow -> flow completion -> image propagation -> sliding-window generator -> compositing."""

    def __init__(self, fix_raft=None, fix_flow_complete=None, model=None, device="cuda", seeds=(1, 2, 3),
                 weights=(None, None, None)):
        self.device = torch.device(device)
        self.fix_raft = fix_raft if fix_raft is not None else RAFT_bi(weights[0], device, seed=seeds[0])
        self.fix_flow_complete = (fix_flow_complete if fix_flow_complete is not None
                                  else RecurrentFlowCompleteNet(weights[1], seed=seeds[1]).to(device))
        self.model = model if model is not None else InpaintGenerator(model_path=weights[2], seed=seeds[2]).to(device)

    def index(self, ids):
        """device int64 index tensor for a frame list, built once per distinct list (the window plan repeats every clip)"""
        cache = self.__dict__.setdefault("_index_cache", {})
        key = tuple(ids)
        if key not in cache:
            cache[key] = torch.tensor(list(ids), dtype=torch.long, device=self.device)
        return cache[key]

    def state_dicts(self):
        return {"raft": self.fix_raft.fix_raft.state_dict(), "rfc": self.fix_flow_complete.state_dict(),
                "gen": self.model.state_dict()}

    # ---- stage 1 (:302-330)
    def compute_flows(self, frames, cfg):
        T, H, W = frames.shape[1], frames.shape[-2], frames.shape[-1]
        clip = cfg.raft_clip_frames
        if clip is None:
            n = (H // 8) * (W // 8)
            pairs = int(8e9 // (5.4 * n * n))                         # 4 pyramid levels ~ 1.34 N^2 floats per pair
            clip = max(raft_clip_len(W), min(T, pairs // 2 + 1))
        ff, bb = [], []
        for s, e in flow_chunks(T, clip):
            f, b = self.fix_raft(frames[:, s:e], iters=cfg.raft_iter)
            ff.append(f)
            bb.append(b)
        return torch.cat(ff, 1), torch.cat(bb, 1)

    # ---- stage 2 (:341-368)
    def complete_flows(self, gt_flows, flow_masks, cfg):
        net, L = self.fix_flow_complete, gt_flows[0].shape[1]
        if L <= cfg.subvideo_length:
            pred, _ = net.forward_bidirect_flow(gt_flows, flow_masks)
            return net.combine_flow(gt_flows, pred, flow_masks)
        pf, pb = [], []
        for s, e, lo, hi in halo_chunks(L, cfg.subvideo_length, 5):
            sub = (gt_flows[0][:, s:e], gt_flows[1][:, s:e])
            pred, _ = net.forward_bidirect_flow(sub, flow_masks[:, s:e + 1])
            pred = net.combine_flow(sub, pred, flow_masks[:, s:e + 1])
            pf.append(pred[0][:, lo:hi])
            pb.append(pred[1][:, lo:hi])
        return torch.cat(pf, 1), torch.cat(pb, 1)

    # ---- stage 3 (:371-404)
    def propagate_images(self, frames, masks_dilated, pred_flows, cfg):
        T = frames.shape[1]
        masked = frames * (1 - masks_dilated)
        sub = min(100, cfg.subvideo_length)
        if T <= sub:
            prop, um = self.model.img_propagation(masked, pred_flows, masks_dilated, "nearest")
            return frames * (1 - masks_dilated) + prop * masks_dilated, um
        uf, umk = [], []
        for s, e, lo, hi in halo_chunks(T, sub, 10):
            prop, um = self.model.img_propagation(masked[:, s:e], (pred_flows[0][:, s:e - 1], pred_flows[1][:, s:e - 1]),
                                                  masks_dilated[:, s:e], "nearest")
            upd = frames[:, s:e] * (1 - masks_dilated[:, s:e]) + prop * masks_dilated[:, s:e]
            uf.append(upd[:, lo:hi])
            umk.append(um[:, lo:hi])
        return torch.cat(uf, 1), torch.cat(umk, 1)

    # ---- stage 4 (:406-452)
    def generate(self, upd_frames, masks_dilated, upd_masks, pred_flows, ori_u8, cfg, windows=None, comp=None,
                 visited=None):
        T = upd_frames.shape[1]
        plan = window_plan(T, cfg)
        comp = torch.zeros_like(ori_u8) if comp is None else comp
        visited = [False] * T if visited is None else visited
        md = masks_dilated[0].contiguous()
        # encoder features depend only on (frame, mask, updated mask): computed once per clip, not once per window
        enc_all = self.model.encode(upd_frames[0], md, upd_masks[0]).permute(0, 2, 3, 1)     # pixel-major rows: cheap frame gather
        todo = [(wi, nb, refs) for wi, (nb, refs) in enumerate(plan) if windows is None or wi in windows]

        um0 = upd_masks[0]

        def job(nb, refs):
            # frame selections as cached device index tensors / slices: indexing with a Python list builds the index on the
            # host and copies it with a blocking cudaMemcpy, which stalls the issuing thread until the stream has drained
            # (measured: generate() blocked the host for the whole clip, so nothing could be queued behind it)
            idx = self.index(nb + refs)
            a, b = nb[0], nb[-1]                                   # neighbour frames are a contiguous range
            return lambda slot: self.model.forward_features(enc_all.index_select(0, idx).permute(0, 3, 1, 2),
                                                            (pred_flows[0][0, a:b], pred_flows[1][0, a:b]),
                                                            md.index_select(0, idx), um0.index_select(0, idx), len(nb), slot=slot)

        def consume(k, pred):
            nb = todo[k][1]
            ops.composite_blend(pred, md, ori_u8, comp, nb, [not visited[i] for i in nb])
            for i in nb:
                visited[i] = True
        self.run_windows([job(nb, refs) for _, nb, refs in todo], consume, cfg, upd_frames.is_cuda)
        return comp

    def run_windows(self, jobs, consume, cfg, cuda=True):
        """Windows are independent given the stage-3 outputs: keep `cfg.windows_in_flight` of them in flight on side streams
        (each slot owns its own captured graph instance); `consume(k, pred)` -- the compositing -- is replayed on the main stream
        in ascending window order because the 1/2-1/2 blend of inference_propainter.py:445-450 is order-dependent.
        jobs[k](slot) launches window k and returns its prediction."""
        nfl = max(1, int(cfg.windows_in_flight)) if cuda else 1
        if nfl == 1:
            for k, jb in enumerate(jobs):
                consume(k, jb(0))
            return
        main = torch.cuda.current_stream()
        if not hasattr(self, "_side_streams") or len(self._side_streams) < nfl:
            self._side_streams = [torch.cuda.Stream(device=self.device) for _ in range(nfl)]
        pending = []

        def drain(keep):
            while len(pending) > keep:
                slot, k, pred = pending.pop(0)
                main.wait_stream(self._side_streams[slot])
                pred.record_stream(main)
                consume(k, pred)
        for k, jb in enumerate(jobs):
            slot = k % nfl
            drain(nfl - 1)
            st = self._side_streams[slot]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                pred = jb(slot)
            pending.append((slot, k, pred))
        drain(0)

    # ---- whole path
    @torch.no_grad()
    def __call__(self, frames_u8, flow_masks, masks_dilated, cfg=None, return_stages=False):
        """frames_u8 uint8 [T,H,W,3] (host or device); masks float {0,1} [1,T,1,H,W].
        Returns composited uint8 frames [T,H,W,3] on the device."""
        cfg = cfg or InferenceConfig()
        dev = self.device
        ori = frames_u8.to(dev, non_blocking=True)
        flow_masks = flow_masks.to(dev, non_blocking=True).float()
        masks_dilated = masks_dilated.to(dev, non_blocking=True).float()
        frames = ops.u8_to_frames(ori).unsqueeze(0)
        gt = self.compute_flows(frames, cfg)
        pred = self.complete_flows(gt, flow_masks, cfg)
        upd_f, upd_m = self.propagate_images(frames, masks_dilated, pred, cfg)
        comp = self.generate(upd_f, masks_dilated, upd_m, pred, ori, cfg)
        if return_stages:
            return comp, {"gt_flows": gt, "pred_flows": pred, "updated_frames": upd_f, "updated_masks": upd_m}
        return comp
