"""Build an experimental variant of the library next to the shipped one:

    python profiles/build_variant.py vmn -DUA_V_MN=1 -DAT_UNMASKED_LOOP=1
    PROPAINTER_B200_LIB=$PWD/propainter_b200/libpropainter_b200_vmn.so python -m pytest tests/test_gpu_ops.py -m gpu -q
    PROPAINTER_B200_LIB=$PWD/propainter_b200/libpropainter_b200_vmn.so python bench.py --no-cpu-baseline

Same sources and flags as __graft_entry__.build() plus the given -D switches.  Compile-time experiments available:
  UA_V_MN=1          attn_umma.cu   row-major V as an MN-major SWIZZLE_128B_BASE32B operand (measured 127/258 us vs 142/322)
  UA_STAGES=3        attn_umma.cu   third K/V stage (needs UA_V_MN=1; untested)
  AT_UNMASKED_LOOP=1 mma_kernels.cu frame-looping kernel for unmasked windows (untested)
The variant .so is git-ignored and travels to the GPU box with gpurun like the shipped one."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

name, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "propainter_b200", f"libpropainter_b200_{name}.so")
cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
       "-std=c++17", "-shared", "-Xcompiler", "-fPIC", *defs, "-o", out] + [os.path.join(g.CSRC, s) for s in g.SOURCES]
subprocess.check_call(cmd)
print(out)
