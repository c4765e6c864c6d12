"""torchrun --nproc-per-node N profiles/dist_check.py : NCCL run of the time-sharded pipeline vs the single-GPU result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from propainter_b200 import synth  # noqa: E402
from propainter_b200.dist import ShardedProPainter  # noqa: E402
from propainter_b200.inference_propainter import InferenceConfig, ProPainterPipeline  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
T, sub = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (33, 10)
u8, fm, md = synth.make_clip(T, 240, 432, mask="ellipse", seed=0)
pipe = ProPainterPipeline(device=f"cuda:{local}")
cfg = InferenceConfig(raft_iter=4, subvideo_length=sub)
sp = ShardedProPainter(pipe)
a = sp(torch.from_numpy(u8), fm, md, cfg, gather=True)
part, ids = sp(torch.from_numpy(u8), fm, md, cfg)                # the product path: every rank keeps its own final frames
torch.cuda.synchronize()
assert torch.equal(part, a[ids])
print(f"rank {dist.get_rank()}: holds {len(ids)} final frames, sent {sum(sp.last_bytes.values()) / 1e6:.1f} MB point to point: "
      + ", ".join(f"{k} {v / 1e6:.1f}" for k, v in sp.last_bytes.items()), flush=True)
if dist.get_rank() == 0:
    b = pipe(torch.from_numpy(u8), fm, md, cfg)
    d = np.abs(a.cpu().numpy().astype(int) - b.cpu().numpy().astype(int))
    print(f"sharded x{dist.get_world_size()} vs single GPU: T={T} max|diff|={d.max()} differing={float((d > 0).mean()):.2e}")
dist.barrier()
dist.destroy_process_group()
