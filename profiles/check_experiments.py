"""Round-2 kick-off: validates and times the experiments that were written at the end of round 1 without GPU access.
Run on the GPU box (optionally with PROPAINTER_B200_LIB pointing at a variant library from build_variant.py):

    python profiles/check_experiments.py

  1. pp_deform_align_batched vs the single-map kernel (must be bit-identical per map)
  2. RecurrentFlowCompleteNet.forward_bidirect_flow with config.RFC_BATCHED on/off: max diff + stage time at the C2 shape
  3. sparse window attention of the loaded library vs the mma.sync baseline (covers UA_V_MN / UA_STAGES / AT_UNMASKED_LOOP
     builds): max diff + time
Nothing here is part of the test-suite: these paths are off by default until this script has passed on hardware."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from propainter_b200 import _lib, config, ops  # noqa: E402
from propainter_b200.model.recurrent_flow_completion import RecurrentFlowCompleteNet  # noqa: E402
from propainter_b200.window_index import window_key_table  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
print("library:", _lib.LIB_PATH)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- 1. batched deform-align
for (n, H, W, Cin, use_flow, mr) in ((2, 30, 54, 256, False, 5.0), (3, 60, 108, 128, True, 3.0)):
    x = torch.randn(n, H, W, Cin + 128, device=dev)[..., :Cin]                       # strided view like the scan buffers
    o = torch.randn(n, H, W, 432, device=dev)
    fl = torch.randn(n, H, W, 2, device=dev) if use_flow else None
    wp, b, ob = torch.randn(9 * Cin, 128, device=dev) * 0.03, torch.randn(128, device=dev), torch.randn(432, device=dev) * 0.1
    out_b = torch.empty(n, H, W, 128, device=dev)
    ops.deform_align(x, o, fl, mr, wp, b, out_b, o_bias=ob)
    out_s = torch.empty_like(out_b)
    for i in range(n):
        ops.deform_align(x[i], o[i], None if fl is None else fl[i], mr, wp, b, out_s[i], o_bias=ob)
    torch.cuda.synchronize()
    tb = timeit(lambda: ops.deform_align(x, o, fl, mr, wp, b, out_b, o_bias=ob)) * 1e3
    ts = timeit(lambda: ops.deform_align(x[0], o[0], None if fl is None else fl[0], mr, wp, b, out_s[0], o_bias=ob)) * 1e3
    print(f"deform batched n={n} {H}x{W} Cin={Cin}: identical={torch.equal(out_b, out_s)} max|d|={(out_b - out_s).abs().max().item():.2e}"
          f"  batched {tb:.1f} us vs one map {ts:.1f} us")

# ---- 2. flow completion, both directions batched vs two streams (C2: 79 flows at 240x432)
net = RecurrentFlowCompleteNet(None, seed=2).to(dev)
flows = (torch.randn(1, 79, 2, 240, 432, device=dev), torch.randn(1, 79, 2, 240, 432, device=dev))
masks = torch.zeros(1, 80, 1, 240, 432, device=dev)
masks[..., 80:160, 150:280] = 1
res = {}
for flag in (False, True):
    config.RFC_BATCHED = flag
    res[flag] = net.forward_bidirect_flow(flows, masks)[0]
    ms = timeit(lambda: net.forward_bidirect_flow(flows, masks), reps=3)
    print(f"RFC_BATCHED={flag}: {ms:.2f} ms per forward_bidirect_flow")
config.RFC_BATCHED = False
for k in (0, 1):
    d = (res[True][k] - res[False][k]).abs().max().item() / res[False][k].abs().max().item()
    print(f"  direction {k}: batched vs two-stream rel max diff {d:.2e}")

# ---- 3. attention of this library build vs the mma.sync baseline
t, H2, W2, C = 18, 20, 36, 512
qkv = torch.randn(t, H2 * W2, 3 * C, device=dev)
pool = torch.randn(t, 45, 2 * C, device=dev)
ktab = torch.from_numpy(window_key_table(H2, W2)).to(dev)
for nm, masked in (("0of16", []), ("5of16", [5, 6, 9, 10, 11]), ("16of16", list(range(16)))):
    flags = torch.zeros(16, dtype=torch.int32, device=dev)
    flags[masked] = 1
    a = ops.sparse_window_attn(qkv, pool, ktab, flags, t, H2 * W2, 0, 2, impl="umma")
    b = ops.sparse_window_attn(qkv, pool, ktab, flags, t, H2 * W2, 0, 2, impl="mma")
    torch.cuda.synchronize()
    err = (a - b).abs().max().item() / b.abs().max().item()
    tu = timeit(lambda: ops.sparse_window_attn(qkv, pool, ktab, flags, t, H2 * W2, 0, 2, impl="umma"), 10) * 1e3
    tm = timeit(lambda: ops.sparse_window_attn(qkv, pool, ktab, flags, t, H2 * W2, 0, 2, impl="mma"), 10) * 1e3
    print(f"attention {nm}: default entry vs mma entry rel max diff {err:.2e}; {tu:.1f} us vs {tm:.1f} us")
