"""Parameter containers that reproduce the reference's ``state_dict`` schema.

The reference nets are loaded with ``load_state_dict(strict=True)`` (model/propainter.py:307-310,
model/recurrent_flow_completion.py:266-269, model/modules/flow_comp_raft.py:18-20), so the drop-in
contract is the exact set of dotted key names, shapes and dtypes.  Instead of mirroring the
reference's ``nn.Module`` class tree we describe each net with a flat ``Schema`` (one line per
layer) and materialise it as a tree of bare ``nn.Module`` nodes whose dotted paths equal the
reference keys.  Forward code looks tensors up by key (``self.P["update_block.gru.convz1.weight"]``).
"""
import math

import torch
import torch.nn as nn


class Schema:
    """Ordered list of (key, shape, dtype, kind, init) entries + module aliases."""

    def __init__(self):
        self.entries = []
        self.aliases = []          # (alias_module_path, target_module_path)

    def add(self, key, shape, kind="param", dtype=torch.float32, init=("zeros",)):
        self.entries.append((key, tuple(shape), dtype, kind, init))

    def conv(self, p, cin, cout, k, groups=1, bias=True, nd=2, gain=1.0):
        ks = (k,) * nd if isinstance(k, int) else tuple(k)
        fan_in = (cin // groups) * math.prod(ks)
        self.add(p + ".weight", (cout, cin // groups) + ks, init=("normal", gain / math.sqrt(fan_in)))
        if bias:
            self.add(p + ".bias", (cout,), init=("normal", 0.02))

    def linear(self, p, cin, cout, gain=1.0):
        self.add(p + ".weight", (cout, cin), init=("normal", gain / math.sqrt(cin)))
        self.add(p + ".bias", (cout,), init=("normal", 0.02))

    def affine(self, p, c):
        self.add(p + ".weight", (c,), init=("normal1", 0.1))
        self.add(p + ".bias", (c,), init=("normal", 0.1))

    def batchnorm(self, p, c):
        self.affine(p, c)
        self.add(p + ".running_mean", (c,), kind="buffer", init=("normal", 0.1))
        self.add(p + ".running_var", (c,), kind="buffer", init=("uniform", 0.8, 1.2))
        self.add(p + ".num_batches_tracked", (), kind="buffer", dtype=torch.int64, init=("zeros",))

    def alias(self, alias_path, target_path):
        self.aliases.append((alias_path, target_path))


class _Node(nn.Module):
    """Bare container; exists only so dotted state_dict paths resolve."""


def _descend(root, parts, create=True):
    m = root
    for name in parts:
        nxt = m._modules.get(name)
        if nxt is None:
            if not create:
                raise KeyError(".".join(parts))
            nxt = _Node()
            m.add_module(name, nxt)
        m = nxt
    return m


def _make_tensor(shape, dtype, init, gen):
    kind = init[0]
    if kind == "zeros":
        return torch.zeros(shape, dtype=dtype)
    if kind == "const":
        return torch.as_tensor(init[1]).to(dtype).reshape(shape).clone()
    if kind == "normal":
        return torch.randn(shape, generator=gen, dtype=torch.float32).mul_(init[1]).to(dtype)
    if kind == "normal1":
        return torch.randn(shape, generator=gen, dtype=torch.float32).mul_(init[1]).add_(1.0).to(dtype)
    if kind == "uniform":
        return torch.rand(shape, generator=gen, dtype=torch.float32).mul_(init[2] - init[1]).add_(init[1]).to(dtype)
    if kind == "normal_mean":
        return torch.randn(shape, generator=gen, dtype=torch.float32).mul_(init[2]).add_(init[1]).to(dtype)
    raise ValueError(kind)


class ParamNet(nn.Module):
    """nn.Module whose parameters/buffers are declared by a Schema.

    Supports .to()/.half()/.eval()/.parameters()/state_dict()/load_state_dict(strict=True) like the
    reference nets.  ``self.P`` is a flat dict key -> live tensor, rebuilt lazily after any
    ``_apply`` (device / dtype move) or ``load_state_dict``; derived (re-packed) weights are cached
    in ``self._packed`` under the same invalidation.
    """

    def __init__(self, schema, seed=None):
        super().__init__()
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        self._keys = []
        for key, shape, dtype, kind, init in schema.entries:
            parts = key.split(".")
            node = _descend(self, parts[:-1])
            t = _make_tensor(shape, dtype, init, gen)
            if kind == "param":
                node.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))
            else:
                node.register_buffer(parts[-1], t)
            self._keys.append(key)
        for alias_path, target_path in schema.aliases:
            tgt = _descend(self, target_path.split("."), create=False)
            ap = alias_path.split(".")
            _descend(self, ap[:-1]).add_module(ap[-1], tgt)
        self._flat = None
        self._packed = {}
        from .graphs import GraphCache
        self.graphs = GraphCache()          # captured stage graphs; dropped whenever the weights change
        self.register_load_state_dict_post_hook(lambda m, _k: m._invalidate())
        self.eval()

    def _invalidate(self):
        self._flat = None
        self._packed = {}
        if hasattr(self, "graphs"):
            self.graphs.clear()

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._invalidate()
        return out

    @property
    def P(self):
        if self._flat is None:
            flat = {}
            for key in self._keys:
                parts = key.split(".")
                node = _descend(self, parts[:-1], create=False)
                t = getattr(node, parts[-1])
                # .half() nets (reference inference_propainter.py:268-270, --fp16): storage may be fp16/bf16, the
                # kernels compute in fp32, so the live view the forward code sees is widened once per dtype move.
                flat[key] = t.float() if t.is_floating_point() and t.dtype != torch.float32 else t
            self._flat = flat
        return self._flat

    def packed(self, name, builder):
        """Cache a derived tensor (re-laid-out / concatenated weights) until params change."""
        if name not in self._packed:
            with torch.no_grad():
                self._packed[name] = builder()
        return self._packed[name]

    @property
    def device(self):
        return next(iter(self.P.values())).device
