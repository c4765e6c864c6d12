"""Per-stage aten-op attribution (with input shapes) of one bench step, CUDA graphs off so that the profiler sees
the eager ops behind every elementwise / copy / cat kernel.  Complements torch_profile_step.py (graphs on, kernels only).
usage: python profiles/torch_profile_ops.py [c2|c1] > gpurun_out/ops_breakdown.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from propainter_b200 import config, ops, synth  # noqa: E402
from propainter_b200.inference_propainter import InferenceConfig, ProPainterPipeline  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

config.CUDA_GRAPHS = False
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
u8, fm, md = synth.make_clip(wl["T"], wl["H"], wl["W"], mask=wl["mask"], seed=0)
pipe = ProPainterPipeline(device="cuda")
cfg = InferenceConfig(raft_iter=wl["raft_iter"], windows_in_flight=1)
u8d, fmd, mdd = torch.from_numpy(u8).cuda(), fm.cuda(), md.cuda()
for _ in range(2):
    pipe(u8d, fmd, mdd, cfg)
torch.cuda.synchronize()


def prof(name, fn):
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as p:
        out = fn()
        torch.cuda.synchronize()
    print(f"\n===== {name}")
    print(p.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=28, max_name_column_width=44,
                                                           max_shapes_column_width=90))
    return out


with torch.no_grad():
    frames = ops.u8_to_frames(u8d).unsqueeze(0)
    gt = prof("1 raft", lambda: pipe.compute_flows(frames, cfg))
    pred = prof("2 flow completion", lambda: pipe.complete_flows(gt, fmd, cfg))
    upd = prof("3 image propagation", lambda: pipe.propagate_images(frames, mdd, pred, cfg))
    prof("4 generator+composite", lambda: pipe.generate(upd[0], mdd, upd[1], pred, u8d, cfg))
