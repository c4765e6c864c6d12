"""Small torch helpers shared by the three nets: pixel-major <-> channels-last views and conv wrappers.

Convolutions / plain Linear layers stay library calls (cuDNN / cuBLAS through torch, SURVEY.md §2b
"keep cuDNN initially"); everything else on the hot path goes through propainter_b200.ops.
"""
import torch
import torch.nn.functional as F

from . import config, ops


def cl(w):
    """conv weight -> channels_last memory format (done once at pack time)."""
    return w.contiguous(memory_format=torch.channels_last)


def as_nchw(pm):
    """pixel-major [n,h,w,c] (dense) -> NCHW-logical channels_last view, no copy."""
    return pm.permute(0, 3, 1, 2)


def as_pm(x):
    """NCHW-logical tensor -> pixel-major [n,h,w,c]; copies only if x is not channels_last."""
    y = x.permute(0, 2, 3, 1)
    return y if y.is_contiguous() else y.contiguous()


def pad_in_channels(w, to):
    """zero-pad a conv weight's input channels (numerically a no-op; keeps pixel rows 16-byte aligned
    and the cuDNN tensor-core path eligible)."""
    if w.shape[1] == to:
        return w
    z = w.new_zeros(w.shape[0], to - w.shape[1], *w.shape[2:])
    return torch.cat([w, z], 1)


_TORCH_ACT = {
    "none": lambda y, s: y,
    "relu": lambda y, s: F.relu_(y),
    "leaky": lambda y, s: F.leaky_relu_(y, s),
    "sigmoid": lambda y, s: torch.sigmoid_(y),
    "tanh": lambda y, s: torch.tanh_(y),
}


def conv(x, wb, stride=1, padding=0, dilation=1, groups=1, act="none", slope=0.0):
    """conv2d + bias + activation.  The conv is cuDNN; bias + activation are one pass of pp_bias_act on the
    channels-last result (cuDNN would launch a separate bias add_, ATen another kernel for the activation).
    Outputs whose channel count is not a multiple of 4 (2/3/126-channel heads) keep the library epilogue."""
    w, b = wb
    if config.FUSED_EPILOGUE and w.shape[0] % 4 == 0:
        y = F.conv2d(x, w, None, stride=stride, padding=padding, dilation=dilation, groups=groups)
        ypm = y.permute(0, 2, 3, 1)
        if ypm.is_contiguous():
            ops.bias_act_(ypm, b, act, slope)
            return y
        return _TORCH_ACT[act](y.add_(b.view(1, -1, 1, 1)), slope)
    y = F.conv2d(x, w, b, stride=stride, padding=padding, dilation=dilation, groups=groups)
    return _TORCH_ACT[act](y, slope)


def up2(x):
    """bilinear x2, align_corners=True, on an NCHW-logical channels_last tensor (the `deconv` blocks)."""
    if x.shape[1] % 4 == 0:
        return as_nchw(ops.upsample2x(as_pm(x)))
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
