"""Round-2 kick-off: validates and times the experiments that were written at the end of round 1 without GPU access.
Run on the GPU box (optionally with PROPAINTER_B200_LIB pointing at a variant library from build_variant.py):

    python profiles/check_experiments.py

  sparse window attention of the loaded library vs the mma.sync baseline (covers UA_V_MN / UA_STAGES / AT_UNMASKED_LOOP
  builds): max diff + time.
(Round-2 outcome of the other two experiments this script used to hold -- gpurun_out exp_default.log: the batched
deform-align entry was not bit-identical per map (2e-6) and RFC_BATCHED was slower (32.3 vs 30.3 ms) than two streams; both
were removed.  UA_V_MN=1 + UA_STAGES=3: 110 / 215 us vs 135 / 312 us -> now the default.)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from propainter_b200 import _lib, ops  # noqa: E402
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
