"""Execution switches (module-level; read at call time).

LINEAR_TF32     plain Linear layers of the transformer (q/k/v, proj, fc1, fc2 = 807 GFLOP per generator call)
                run as TF32 tensor-core GEMMs (fp32 accumulate) instead of cuBLAS' fp32 SIMT sgemm.  Same
                precision class as the reference's own CUDA convs (cuDNN TF32 is torch's default) and as our
                attention / deform-align kernels.  Set False for fp32-exact library GEMMs (tests do, to
                isolate kernel error).
CUDNN_BENCHMARK let cuDNN autotune conv algorithms during graph warm-up.
CUDA_GRAPHS     replay each stage as a captured CUDA graph per shape signature (propainter_b200/graphs.py).
FUSED_EPILOGUE  conv bias + activation through pp_bias_act (one pass) instead of cuDNN's bias add_ + ATen activation.
RFC_BATCHED     run the forward-flow and backward-flow nets of forward_bidirect_flow as one batch of two (every kernel of
                the scans serves both; pp_deform_align_batched) instead of two concurrent streams.  The scans' kernels
                occupy a fraction of the GPU, so batching should cost one scan instead of ~1.5 (two streams slow each
                other down).  Off until it has been validated and timed on hardware.
AUTOTUNE        time numerically equivalent plans of a step once per shape during warm-up and keep the faster
                (propainter_b200/autotune.py): grouped conv vs per-group dense convs, conv + pp_bias_act vs cuDNN's fused
                conv-bias-ReLU.
"""
import contextlib

import torch

LINEAR_TF32 = True
CUDNN_BENCHMARK = True
CUDA_GRAPHS = True
FUSED_EPILOGUE = True
AUTOTUNE = True
RFC_BATCHED = False


@contextlib.contextmanager
def linear_precision():
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = bool(LINEAR_TF32) or prev
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


@contextlib.contextmanager
def cudnn_autotune():
    prev = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = bool(CUDNN_BENCHMARK) or prev
    try:
        yield
    finally:
        torch.backends.cudnn.benchmark = prev
