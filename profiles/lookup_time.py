"""CUDA-event timing of the RAFT lookup at whole-clip batch (158 pairs, 30x54): TMA-staged kernel vs plain-load baseline."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from propainter_b200 import ops  # noqa: E402

dev = "cuda"
B, h, w = 158, 30, 54
fmap = torch.randn(80, h * w, 256, device=dev)
a = torch.arange(79, device=dev, dtype=torch.int32)
lv = ops.corr_alloc(B, h, w, dev)
ops.corr_build(fmap, torch.cat([a, a + 1]), torch.cat([a + 1, a]), lv, h, w)
ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
c = (torch.stack([xs, ys], -1).float()[None] + torch.randn(B, h, w, 2, device=dev) * 3).contiguous()
out = torch.empty(B, h, w, 324, device=dev)
fl = torch.empty(64 * 1024 * 1024, device=dev)
alg = B * (h * w * 4 * 100 * 4 + h * w * 324 * 4 + h * w * 8)
for tma in (True, False):
    ts = []
    for _ in range(10):
        fl.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.corr_lookup(lv, c, out, tma=tma)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    med = statistics.median(ts[2:])
    print(f"{'tma' if tma else 'ldg'}: {med:8.1f} us  {alg / med / 1e3:8.1f} GB/s algorithmic ({alg / 1e6:.0f} MB)")
