"""Execution switches (module-level; read at call time).

LINEAR_TF32     plain Linear layers of the transformer (q/k/v, proj, fc1, fc2 = 807 GFLOP per generator call)
                run as TF32 tensor-core GEMMs (fp32 accumulate) instead of cuBLAS' fp32 SIMT sgemm.  Same
                precision class as the reference's own CUDA convs (cuDNN TF32 is torch's default) and as our
                attention / deform-align kernels.  Set False for fp32-exact library GEMMs (tests do, to
                isolate kernel error).
CUDNN_BENCHMARK let cuDNN autotune conv algorithms during graph warm-up.
CUDA_GRAPHS     replay each stage as a captured CUDA graph per shape signature (propainter_b200/graphs.py).
FUSED_EPILOGUE  conv bias + activation through pp_bias_act (one pass) instead of cuDNN's bias add_ + ATen activation.
UMMA_CONV       the convolutions of the two recurrent propagation scans (offset nets, backbones, deformable-conv GEMM) run on
                the tcgen05 implicit-GEMM kernel pp_conv2d_umma (TF32 products, fused bias / activation / residual / concat
                epilogue) instead of cuDNN + pp_bias_act + the mma.sync deform kernel.  True / False force one plan;
                "hybrid" keeps the library convs and replaces only the deformable conv by pp_deform_gather + a 1x1
                pp_conv2d_umma GEMM; "hoisted" additionally convolves the step-independent input channels of
                conv_offset.0 / backbone.0 once per scan (library convs, pp_bias_act_pre); "auto" (default) times the plans of a scan once per shape during graph warm-up
                (autotune.pick) and replays the fastest.  Environment: PP_UMMA_CONV=1|0|hybrid|hoisted|auto.
SCAN_PRIORITY   capture the recurrent propagation scans as high-priority branches of their stage graphs (graphs.high_priority).
                Environment: PP_SCAN_PRIORITY=0|1.
GRAPH_MAX_INPUT_BYTES  stage calls whose inputs exceed this run eagerly instead of as a captured graph (memory: a capture keeps
                its whole working set alive in a private pool).
AUTOTUNE        time numerically equivalent plans of a step once per shape during warm-up and keep the faster
                (propainter_b200/autotune.py): grouped conv vs per-group dense convs, conv + pp_bias_act vs cuDNN's fused
                conv-bias-ReLU.
"""
import contextlib
import os

import torch

LINEAR_TF32 = True
CUDNN_BENCHMARK = True
CUDA_GRAPHS = True
FUSED_EPILOGUE = True
AUTOTUNE = True
GRAPH_MAX_INPUT_BYTES = 512 << 20
SCAN_PRIORITY = os.environ.get("PP_SCAN_PRIORITY", "1") != "0"
_u = os.environ.get("PP_UMMA_CONV", "auto")
UMMA_CONV = _u if _u in ("auto", "hybrid", "hoisted") else (_u != "0")


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
