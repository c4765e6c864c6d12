"""``RAFT_bi``: bidirectional flow between consecutive frames.

Drop-in for model/modules/flow_comp_raft.py:10-55 of the reference (same constructor, same
``forward(gt_local_frames, iters) -> (flows_forward, flows_backward)``); the training losses in the
rest of that file (:58-264) are outside the inference hot path.
"""
import torch
import torch.nn as nn

from ...RAFT import RAFT


def initialize_RAFT(model_path="weights/raft-things.pth", device="cuda", seed=None):
    """flow_comp_raft.py:10-24.  The released checkpoint was saved through nn.DataParallel, so its keys carry a
    ``module.`` prefix (:18-20); it is loaded strict.  ``model_path=None`` gives a seeded random init (there are
    no weights in the build environment)."""
    model = RAFT(seed=seed)
    if model_path is not None:
        sd = torch.load(model_path, map_location="cpu")
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
        model.load_state_dict(sd, strict=True)
    return model.to(device)


class RAFT_bi(nn.Module):
    def __init__(self, model_path="weights/raft-things.pth", device="cuda", seed=None):
        super().__init__()
        self.fix_raft = initialize_RAFT(model_path, device=device, seed=seed)
        for p in self.fix_raft.parameters():
            p.requires_grad = False
        self.eval()

    @torch.no_grad()
    def forward(self, gt_local_frames, iters=20):
        """gt_local_frames [b,l,3,h,w] -> (fwd, bwd) each [b,l-1,2,h,w]  (flow_comp_raft.py:39-55)."""
        b, l, c, h, w = gt_local_frames.shape
        fw, bw = [], []
        for i in range(b):
            f, g = self.fix_raft.flows_bidirectional(gt_local_frames[i], iters=iters)
            fw.append(f)
            bw.append(g)
        return torch.stack(fw, 0).view(b, l - 1, 2, h, w), torch.stack(bw, 0).view(b, l - 1, 2, h, w)
