"""Per-CTA cycle attribution of k_conv_umma (needs the CV_PROFILE variant of the library):

    python profiles/build_variant.py prof -DCV_PROFILE=1
    PROPAINTER_B200_LIB=$PWD/propainter_b200/libpropainter_b200_prof.so python profiles/conv_prof.py

Slots written by the kernel (clock64 cycles): producer start / wait on a_empty / wait on b_empty / producer end;
MMA thread start / wait on a_full / wait on b_full / first A tile seen / MMA loop end; epilogue saw acc_full; CTA end."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from propainter_b200 import _lib, ops  # noqa: E402

dev = "cuda"
L = _lib.lib()
setbuf = L.pp_conv_profile_buffer
setbuf.argtypes, setbuf.restype = [ctypes.c_void_p], None
buf = torch.zeros(4096 * 16, dtype=torch.int64, device=dev)
setbuf(ctypes.c_void_p(buf.data_ptr()))


def case(name, n, H, W, Cin, Cout, KH, KW, bn=0, tile_m=0):
    x = torch.randn(n, H, W, Cin, device=dev)
    w = torch.randn(Cout, Cin, KH, KW, device=dev) * 0.05
    wp = ops.pack_conv_weight(w)
    out = torch.empty(n, H, W, Cout, device=dev)
    for _ in range(3):
        buf.zero_()
        ops.conv_umma([x], wp, KH, KW, Cout, out=out, bn=bn, tile_m=tile_m)
        torch.cuda.synchronize()
    b = buf.view(-1, 16).cpu()
    b = b[b[:, 10] != 0]
    t0 = b[:, 0].min()
    f = lambda v: f"{v.float().mean().item():9.0f}"
    print(f"{name}: {b.shape[0]} CTAs | CTA lifetime {f(b[:, 10] - b[:, 0])} cyc | kernel span {int((b[:, 10].max() - t0))} cyc\n"
          f"   producer: wait a_empty {f(b[:, 1])}  wait b_empty {f(b[:, 2])}  done at +{f(b[:, 3] - b[:, 0])}\n"
          f"   mma     : wait a_full  {f(b[:, 5])}  wait b_full  {f(b[:, 6])}  first A at +{f(b[:, 7] - b[:, 4])}  loop end +{f(b[:, 8] - b[:, 4])}\n"
          f"   epilogue: acc_full at +{f(b[:, 9] - b[:, 0])}  staged +{f(b[:, 12] - b[:, 9])}  stored +{f(b[:, 14] - b[:, 9])}  end(sync) +{f(b[:, 10] - b[:, 9])}", flush=True)


case("3x3 rfc 128->128 auto", 1, 30, 54, 128, 128, 3, 3)
if os.environ.get("PROF_QUICK"):
    sys.exit(0)
case("3x3 gen 128->128 bn64", 1, 60, 108, 128, 128, 3, 3, 64)
case("3x3 gen 128->128 bn128", 1, 60, 108, 128, 128, 3, 3, 128)
case("1x1 gen K=1152 auto", 1, 60, 108, 1152, 128, 1, 1)
case("3x3 gen 128->128 M=64 bn128", 1, 60, 108, 128, 128, 3, 3, 128, 64)
case("3x3 gen 128->128 M=64 bn64", 1, 60, 108, 128, 128, 3, 3, 64, 64)
case("3x3 rfc 128->128 M=64 bn64", 1, 30, 54, 128, 128, 3, 3, 64, 64)
case("3x3 rfc 128->128 M=64 bn32", 1, 30, 54, 128, 128, 3, 3, 32, 64)
case("1x1 tiny K=128", 1, 16, 8, 128, 32, 1, 1)
case("3x3 tiny 32->32", 1, 16, 8, 32, 32, 3, 3)
