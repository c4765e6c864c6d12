"""Host issue time per stage (no syncs between stages) and per GraphCache call."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from propainter_b200 import graphs, ops, synth  # noqa: E402
from propainter_b200.inference_propainter import InferenceConfig, ProPainterPipeline  # noqa: E402

u8, fm, md = synth.make_clip(80, 240, 432, mask="ellipse", seed=0)
dev = "cuda"
x, fm, md = torch.from_numpy(u8).to(dev), fm.to(dev), md.to(dev)
cfg = InferenceConfig()
pipe = ProPainterPipeline(device=dev)
for _ in range(3):
    pipe(x, fm, md, cfg)
torch.cuda.synchronize()

# instrument GraphCache.__call__ and CUDAGraph.replay
acc = {}
orig_call = graphs.GraphCache.__call__
orig_replay = torch.cuda.CUDAGraph.replay


def timed_call(self, key, fn, *inputs):
    t0 = time.perf_counter()
    r = orig_call(self, key, fn, *inputs)
    k = key if isinstance(key, str) else key[0]
    a = acc.setdefault(("cache", k), [0, 0.0])
    a[0] += 1
    a[1] += time.perf_counter() - t0
    return r


def timed_replay(self):
    t0 = time.perf_counter()
    orig_replay(self)
    a = acc.setdefault(("replay", ""), [0, 0.0])
    a[0] += 1
    a[1] += time.perf_counter() - t0


graphs.GraphCache.__call__ = timed_call
torch.cuda.CUDAGraph.replay = timed_replay
for rep in range(2):
    acc.clear()
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    frames = ops.u8_to_frames(x).unsqueeze(0)
    gt = pipe.compute_flows(frames, cfg); t.append(time.perf_counter())
    pred = pipe.complete_flows(gt, fm, cfg); t.append(time.perf_counter())
    uf, um = pipe.propagate_images(frames, md, pred, cfg); t.append(time.perf_counter())
    comp = pipe.generate(uf, md, um, pred, x, cfg); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    vals = [1e3 * (t[i + 1] - t[i]) for i in range(5)] + [1e3 * (t[-1] - t[0])]
    print("host issue ms: raft %.1f  complete %.1f  imgprop %.1f  generate %.1f  | drain %.1f  total %.1f" % tuple(vals), flush=True)
    for k, (n, s) in sorted(acc.items()):
        print(f"   {k}: {n} calls, {1e3 * s:.1f} ms host", flush=True)
