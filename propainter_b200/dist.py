"""One clip across several GPUs (one process per GPU, torch.distributed; NCCL over NVLink on B200, gloo in the CPU tests).

The reference's inference is single-device; what it *does* have is a decomposition of every stage
into independent units with recompute halos (inference_propainter.py:302-319, :342-364, :373-398, :417-452).  Those
units are the shard boundary here, so the sharded result is the same math per unit as the single-GPU run:

  stage 1  RAFT          unit = contiguous range of frame pairs (+1 halo frame)        -> all ranks get all flows
  stage 2  completion    unit = sub-video of `subvideo_length` flows (+5-flow input halo)   (broadcast from the owner)
  stage 3  image prop.   unit = sub-video of min(100, subvideo_length) frames (+10 halo)
  stage 4  generator     unit = sliding window; rank r owns a contiguous run of windows
  merge                  the 1/2-1/2 blend (:445-450) is order dependent, so compositing is replayed in ascending window
                         order: rank r composites after receiving the frames its first windows share with rank r-1
                         (<= 11 uint8 frames, point-to-point), then the final frames are gathered on every rank.

Only the results of stages 1-3 (flows, propagated frames, masks) and the seam frames cross ranks; there is no collective
inside a stage.  For an 80-frame clip stages 2-3 are a single unit each and therefore do not speed up.
"""
import torch
import torch.distributed as dist

from . import ops
from .inference_propainter import InferenceConfig, flow_chunks, halo_chunks, raft_clip_len, window_plan


def split_range(n, parts):
    """[lo, hi) of `n` items for each of `parts` owners, contiguous and as even as possible."""
    base, rem = divmod(n, parts)
    out, lo = [], 0
    for r in range(parts):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def window_owner(n_windows, world):
    """owner rank of every sliding window (contiguous runs, ascending)."""
    owner = [0] * n_windows
    for r, (lo, hi) in enumerate(split_range(n_windows, world)):
        for i in range(lo, hi):
            owner[i] = r
    return owner


def final_frame_owner(plan, owner):
    """rank holding the final value of each frame = owner of the last window that visits it."""
    last = {}
    for wi, (nb, _) in enumerate(plan):
        for f in nb:
            last[f] = owner[wi]
    return last


class ShardedProPainter:
    def __init__(self, pipe, group=None):
        self.pipe, self.group = pipe, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def _bcast(self, t, src):
        dist.broadcast(t, src=src, group=self.group)
        return t

    def _gather_units(self, units, compute, shape_of, like):
        """units: list of ids; compute(u) -> tensor on the owner (u % world); every rank ends up with all results."""
        res = []
        for i, u in enumerate(units):
            owner = i % self.world
            buf = compute(u).contiguous() if owner == self.rank else like.new_empty(shape_of(u))
            res.append(self._bcast(buf, owner))
        return res

    @torch.no_grad()
    def __call__(self, frames_u8, flow_masks, masks_dilated, cfg=None):
        cfg = cfg or InferenceConfig()
        pipe, dev = self.pipe, self.pipe.device
        ori = frames_u8.to(dev)
        flow_masks, masks_dilated = flow_masks.to(dev).float(), masks_dilated.to(dev).float()
        frames = ops.u8_to_frames(ori).unsqueeze(0)
        T, H, W = frames.shape[1], frames.shape[-2], frames.shape[-1]

        # ---- stage 1: frame pairs split evenly; each rank runs RAFT on its range (+1 halo frame)
        pr = [(lo, hi) for lo, hi in split_range(T - 1, self.world) if hi > lo]
        fw, bw = [], []
        for i, (lo, hi) in enumerate(pr):
            owner = i % self.world
            if owner == self.rank:
                f, b = pipe.fix_raft(frames[:, lo:hi + 1], iters=cfg.raft_iter)
                both = torch.stack([f[0], b[0]], 0).contiguous()
            else:
                both = frames.new_empty(2, hi - lo, 2, H, W)
            self._bcast(both, owner)
            fw.append(both[0])
            bw.append(both[1])
        gt = (torch.cat(fw, 0).unsqueeze(0), torch.cat(bw, 0).unsqueeze(0))

        # ---- stage 2: sub-videos of flows (halo chunks of inference_propainter.py:342-364), one owner each
        L = T - 1
        net = pipe.fix_flow_complete
        units = halo_chunks(L, cfg.subvideo_length, 5) if L > cfg.subvideo_length else [(0, L, 0, L)]
        pf, pb = [], []
        for i, (s, e, lo, hi) in enumerate(units):
            owner = i % self.world
            if owner == self.rank:
                sub = (gt[0][:, s:e], gt[1][:, s:e])
                pred, _ = net.forward_bidirect_flow(sub, flow_masks[:, s:e + 1])
                pred = net.combine_flow(sub, pred, flow_masks[:, s:e + 1])
                both = torch.stack([pred[0][0, lo:hi], pred[1][0, lo:hi]], 0).contiguous()
            else:
                both = frames.new_empty(2, hi - lo, 2, H, W)
            self._bcast(both, owner)
            pf.append(both[0])
            pb.append(both[1])
        pred_flows = (torch.cat(pf, 0).unsqueeze(0), torch.cat(pb, 0).unsqueeze(0))

        # ---- stage 3: image propagation units (:373-398)
        sub_len = min(100, cfg.subvideo_length)
        masked = frames * (1 - masks_dilated)
        units = halo_chunks(T, sub_len, 10) if T > sub_len else [(0, T, 0, T)]
        uf, um = [], []
        for i, (s, e, lo, hi) in enumerate(units):
            owner = i % self.world
            if owner == self.rank:
                prop, m = pipe.model.img_propagation(masked[:, s:e], (pred_flows[0][:, s:e - 1], pred_flows[1][:, s:e - 1]),
                                                     masks_dilated[:, s:e], "nearest")
                upd = frames[:, s:e] * (1 - masks_dilated[:, s:e]) + prop * masks_dilated[:, s:e]
                both = torch.cat([upd[0, lo:hi], m[0, lo:hi]], 1).contiguous()          # [n, 3+1, H, W]
            else:
                both = frames.new_empty(hi - lo, 4, H, W)
            self._bcast(both, owner)
            uf.append(both[:, :3])
            um.append(both[:, 3:])
        upd_f, upd_m = torch.cat(uf, 0).unsqueeze(0), torch.cat(um, 0).unsqueeze(0)

        # ---- stage 4: contiguous runs of windows; ordered compositing across the seams
        plan = window_plan(T, cfg)
        owner = window_owner(len(plan), self.world)
        mine = [wi for wi in range(len(plan)) if owner[wi] == self.rank]
        comp = torch.zeros_like(ori)
        visited = [False] * T
        md = masks_dilated[0].contiguous()
        preds = {}
        if mine:
            enc_ids = sorted({f for wi in mine for f in plan[wi][0] + plan[wi][1]})
            pos = {f: i for i, f in enumerate(enc_ids)}
            enc = pipe.model.encode(upd_f[0, enc_ids], md[enc_ids], upd_m[0, enc_ids]).permute(0, 2, 3, 1)   # pixel-major
            for wi in mine:
                nb, refs = plan[wi]
                ids = nb + refs
                sel = [pos[f] for f in ids]
                preds[wi] = pipe.model.forward_features(enc[sel].permute(0, 3, 1, 2), (pred_flows[0][0, nb[:-1]], pred_flows[1][0, nb[:-1]]),
                                                        md[ids], upd_m[0, ids], len(nb))
        # seam state from the previous non-empty rank: every frame visited by an earlier window
        earlier = sorted({f for wi in range(len(plan)) if owner[wi] < self.rank for f in plan[wi][0]})
        need = sorted({f for wi in mine for f in plan[wi][0]} & set(earlier))
        prev = max([owner[wi] for wi in range(len(plan)) if owner[wi] < self.rank], default=None)
        if need and prev is not None:
            buf = comp.new_empty(len(need), H, W, 3)
            dist.recv(buf, src=prev, group=self.group)
            comp[need] = buf
            for f in need:
                visited[f] = True
        for wi in mine:
            nb = plan[wi][0]
            ops.composite_blend(preds[wi], md, ori, comp, nb, [not visited[i] for i in nb])
            for i in nb:
                visited[i] = True
        nxt = min([owner[wi] for wi in range(len(plan)) if owner[wi] > self.rank], default=None)
        if nxt is not None and mine:
            later = {f for wi in range(len(plan)) if owner[wi] == nxt for f in plan[wi][0]}
            done = {f for wi in range(len(plan)) if owner[wi] <= self.rank for f in plan[wi][0]}
            send = sorted(later & done)
            if send:
                dist.send(comp[send].contiguous(), dst=nxt, group=self.group)

        # ---- merge: every rank contributes the frames whose final value it holds
        fin = final_frame_owner(plan, owner)
        allc = [torch.empty_like(comp) for _ in range(self.world)]
        dist.all_gather(allc, comp.contiguous(), group=self.group)
        out = torch.empty_like(comp)
        for f in range(T):
            out[f] = allc[fin[f]][f]
        return out
