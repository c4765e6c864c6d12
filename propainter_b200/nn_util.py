"""Small torch helpers shared by the three nets: pixel-major <-> channels-last views and conv wrappers.

Convolutions / plain Linear layers stay library calls (cuDNN / cuBLAS through torch, SURVEY.md §2b
"keep cuDNN initially"); everything else on the hot path goes through propainter_b200.ops.
"""
import torch
import torch.nn.functional as F


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


def conv(x, wb, stride=1, padding=0, dilation=1, groups=1):
    return F.conv2d(x, wb[0], wb[1], stride=stride, padding=padding, dilation=dilation, groups=groups)
