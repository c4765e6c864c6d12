"""Kernel-time breakdown of one bench step via torch.profiler (CUPTI); complements the ncu captures.
usage: python profiles/torch_profile_step.py [c2|c1] > gpurun_out/breakdown.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from propainter_b200 import synth  # noqa: E402
from propainter_b200.inference_propainter import InferenceConfig, ProPainterPipeline  # noqa: E402

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
u8, fm, md = synth.make_clip(wl["T"], wl["H"], wl["W"], mask=wl["mask"], seed=0)
pipe = ProPainterPipeline(device="cuda")
cfg = InferenceConfig(raft_iter=wl["raft_iter"])
u8d, fmd, mdd = torch.from_numpy(u8).cuda(), fm.cuda(), md.cuda()
for _ in range(2):
    pipe(u8d, fmd, mdd, cfg)
torch.cuda.synchronize()
stage_ms = {}
frames = None
with torch.no_grad():
    from propainter_b200 import ops
    def timed(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        stage_ms[name] = e0.elapsed_time(e1)
        return out
    frames = ops.u8_to_frames(u8d).unsqueeze(0)
    gt = timed("1 raft", lambda: pipe.compute_flows(frames, cfg))
    pred = timed("2 flow completion", lambda: pipe.complete_flows(gt, fmd, cfg))
    upd = timed("3 image propagation", lambda: pipe.propagate_images(frames, mdd, pred, cfg))
    timed("4 generator+composite", lambda: pipe.generate(upd[0], mdd, upd[1], pred, u8d, cfg))
print("stage ms:", {k: round(v, 2) for k, v in stage_ms.items()})
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pipe(u8d, fmd, mdd, cfg)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))
