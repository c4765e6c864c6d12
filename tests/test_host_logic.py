"""CPU tests: state_dict schema vs the reference manifest, host-side scheduling vs the oracle's restated
driver loop, attention index tables vs the roll/partition construction, C-ABI exports."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_schema_matches_reference_manifest():
    from propainter_b200 import schemas
    from propainter_b200._params import ParamNet
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_manifest.json")))
    for name, sch in (("raft", schemas.raft_schema()), ("rfc", schemas.rfc_schema()), ("gen", schemas.generator_schema())):
        sd = ParamNet(sch, seed=0).state_dict()
        mine = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()}
        assert mine == man[name], name


def test_param_net_roundtrip_and_half():
    from propainter_b200 import schemas
    from propainter_b200._params import ParamNet
    a, b = ParamNet(schemas.rfc_schema(), seed=1), ParamNet(schemas.rfc_schema(), seed=2)
    k = "decoder1.0.weight"
    assert not torch.equal(a.P[k], b.P[k])
    b.load_state_dict(a.state_dict(), strict=True)
    assert torch.equal(a.P[k], b.P[k])
    with pytest.raises(RuntimeError):
        b.load_state_dict({k: a.P[k]}, strict=True)                 # strict like the reference loaders
    b.half()
    # .half() halves the storage (state_dict, like the reference's --fp16 nets); the view the kernels read stays fp32
    assert b.state_dict()[k].dtype == torch.float16 and b.P[k].dtype == torch.float32
    assert torch.equal(b.P[k], b.state_dict()[k].float()) and all(not p.requires_grad for p in b.parameters())
    raft = ParamNet(schemas.raft_schema(), seed=0)                   # shared norm3 / downsample.1 module
    assert raft.P["cnet.layer2.0.norm3.weight"] is raft.state_dict(keep_vars=True)["cnet.layer2.0.downsample.1.weight"]


def test_rolled_valid_index_and_key_table():
    from oracle import generator_ref
    from propainter_b200.window_index import rolled_valid_index, window_key_table
    assert np.array_equal(rolled_valid_index((5, 9)), generator_ref._rolled_valid_index().numpy())
    for (H2, W2) in ((20, 36), (15, 18), (5, 9)):
        tab = window_key_table(H2, W2)
        # independent construction through torch.roll + window_partition on a token-id tensor
        ids = torch.arange(H2 * W2).view(1, 1, H2, W2, 1).float()
        own = generator_ref._windows(ids.expand(1, 1, H2, W2, 4).contiguous(), 4)[0, :, 0, 0, :, 0]
        e = (3, 5)
        rolled = [generator_ref._windows(torch.roll(ids, s, (2, 3)).expand(1, 1, H2, W2, 4).contiguous(), 4)[0, :, 0, 0, :, 0]
                  for s in ((-e[0], -e[1]), (-e[0], e[1]), (e[0], -e[1]), (e[0], e[1]))]
        ref = torch.cat([own, torch.cat(rolled, 1)[:, generator_ref._rolled_valid_index()]], 1).long().numpy()
        assert tab.shape == (H2 // 5 * (W2 // 9), 45 + 148) and np.array_equal(tab, ref)


def test_scheduling_matches_oracle_driver():
    from oracle import pipeline_ref
    from propainter_b200.inference_propainter import (InferenceConfig, flow_chunks, get_ref_index, halo_chunks,
                                                      raft_clip_len, window_plan)
    for T in (1, 6, 8, 11, 80, 81, 95, 100, 101, 170, 300):
        for sub in (80, 40):
            cfg = InferenceConfig(subvideo_length=sub)
            assert window_plan(T, cfg) == pipeline_ref.window_plan(T, 10, 10, sub)
    for mid in range(0, 200, 5):
        nb = list(range(max(0, mid - 5), min(200, mid + 6)))
        for rn in (-1, 8, 4):
            assert get_ref_index(mid, nb, 200, 10, rn) == pipeline_ref.get_ref_index(mid, nb, 200, 10, rn)
    for wdt in (432, 640, 641, 720, 1280, 1920):
        assert raft_clip_len(wdt) == pipeline_ref.raft_clip_len(wdt)
    # every consecutive pair is produced exactly once
    for T, clip in ((80, 12), (13, 12), (12, 12), (5, 12), (300, 4)):
        pairs = []
        for s, e in flow_chunks(T, clip):
            pairs += list(range(s, e - 1))
        assert pairs == list(range(T - 1))
    # halo chunks tile [0,L) exactly
    for L, sub, pad in ((79, 80, 5), (299, 80, 5), (300, 80, 10), (1000, 100, 10), (81, 80, 5)):
        kept = []
        for s, e, lo, hi in halo_chunks(L, sub, pad):
            assert 0 <= s <= e <= L and e - s <= sub + 2 * pad
            kept += list(range(s + lo, s + hi))
        assert kept == list(range(L))


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports each symbol include/propainter_b200.h declares (no compute)."""
    import __graft_entry__ as g
    g.build()
    from propainter_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "propainter_b200.h")).read()
    declared = set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        getattr(handle, name)
    assert _lib.lib().pp_abi_version() == 2
    assert _lib.lib().pp_error_string(-3) == b"workspace too small"
    assert _lib.lib().pp_img_prop_scan_workspace_bytes(3, 4, 5) == 3 * 4 * 4 * 5 * 4


def test_ctypes_table_matches_header_prototypes():
    """Every prototype of include/propainter_b200.h, parameter by parameter, against the argtypes the Python side binds
    (`_lib.SIGNATURES`): a transposed or missing argument would otherwise only show up as garbage on the GPU.  The .cu
    files include the same header, so the compiler already ties the prototypes to the definitions."""
    from propainter_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "propainter_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", " ", hdr)
    protos = re.findall(r"\b([A-Za-z_][\w \*]*?)\b(pp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr)
    assert len(protos) == len(_lib.SIGNATURES)

    def ctype(decl):
        decl = decl.strip()
        if "*" in decl:
            for struct in ("PPAttnParams", "PPWindowIds", "PPConvParams"):
                if struct in decl:
                    return ctypes.POINTER(getattr(_lib, struct))
            return ctypes.c_char_p if decl.replace("const", "").strip().startswith("char") and "(" not in decl and decl.count(" ") <= 1 else ctypes.c_void_p
        base = re.sub(r"\b(const|unsigned)\b", "", decl).split()
        kind = base[0]
        return {"int": ctypes.c_int, "long": ctypes.c_long, "float": ctypes.c_float, "size_t": ctypes.c_size_t,
                "cudaStream_t": ctypes.c_void_p, "PPLevels": None}[kind]

    for ret, name, params in protos:
        res, args = _lib.SIGNATURES[name]
        want = [] if params.strip() in ("", "void") else [ctype(a) for a in params.split(",")]
        assert len(want) == len(args), (name, params)
        for i, (w, a) in enumerate(zip(want, args)):
            if w is ctypes.c_char_p:
                w = ctypes.c_void_p
            assert w is a or (w is ctypes.c_void_p and a is ctypes.c_void_p), (name, i, params.split(",")[i].strip(), a)
        ret = ret.strip()
        want_res = ctypes.c_char_p if "char" in ret else ctype(ret + " x")
        assert want_res is res, (name, ret, res)


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: wrappers raise instead of computing on the host."""
    from propainter_b200 import ops
    with pytest.raises(RuntimeError):
        ops.u8_to_frames(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))
    with pytest.raises(RuntimeError):
        ops.img_prop_scan(torch.zeros(2, 3, 8, 8), torch.zeros(1, 2, 8, 8), torch.zeros(1, 2, 8, 8), torch.zeros(2, 1, 8, 8))
