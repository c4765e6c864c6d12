"""One C2 clip between cudaProfilerStart / Stop after three warm-up clips, for an in-situ ncu launch list:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
        -k regex:'k_conv_umma|k_deform_gather|k_sparse_attn_umma|k_attn_unmasked_frames|k_corr_lookup_tma|k_corr_build|k_flow_warp' \
        --csv --log-file gpurun_out/r2_launches_own_kernels.csv python profiles/ncu_bench_step.py

(every graph kernel node of a clip is ~6.6 k launches = ~18 GPU-minutes under ncu, hence the name filter)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from propainter_b200 import synth  # noqa: E402
from propainter_b200.inference_propainter import InferenceConfig, ProPainterPipeline  # noqa: E402

wl = WORKLOADS["c2"]
u8, fm, md = synth.make_clip(wl["T"], wl["H"], wl["W"], mask=wl["mask"], seed=0)
pipe = ProPainterPipeline(device="cuda")
cfg = InferenceConfig(raft_iter=wl["raft_iter"])
u8d, fmd, mdd = torch.from_numpy(u8).cuda(), fm.cuda(), md.cuda()
for _ in range(3):
    pipe(u8d, fmd, mdd, cfg)
torch.cuda.synchronize()
torch.cuda.profiler.start()
pipe(u8d, fmd, mdd, cfg)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
