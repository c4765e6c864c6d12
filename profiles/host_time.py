"""How long does the host take to ISSUE one clip (pipe() returning without a sync) vs the GPU to finish it?
And: two pipelines driven from two Python threads on two streams -- does throughput scale?"""
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from propainter_b200 import synth  # noqa: E402
from propainter_b200.inference_propainter import InferenceConfig, ProPainterPipeline  # noqa: E402

u8, fm, md = synth.make_clip(80, 240, 432, mask="ellipse", seed=0)
dev = "cuda"
x, fm, md = torch.from_numpy(u8).to(dev), fm.to(dev), md.to(dev)
cfg = InferenceConfig()
pipes = [ProPainterPipeline(device=dev) for _ in range(2)]
for p in pipes:
    for _ in range(3):
        p(x, fm, md, cfg)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    pipes[0](x, fm, md, cfg)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"single clip: host issue {1e3 * (t1 - t0):.1f} ms, until GPU done {1e3 * (t2 - t0):.1f} ms", flush=True)


def worker(k, n):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(n):
            pipes[k](x, fm, md, cfg)
    st.synchronize()


for nthreads in (1, 2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ths = [threading.Thread(target=worker, args=(k, 6)) for k in range(nthreads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{nthreads} thread(s) x 6 clips: {1e3 * dt:.0f} ms -> {80 * 6 * nthreads / dt:.1f} frames/s", flush=True)
