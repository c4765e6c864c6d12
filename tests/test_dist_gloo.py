"""world_size-2 `gloo` test of the N>1 host logic (propainter_b200/dist.py) on CPU: the time-sharded pipeline must
reproduce the single-process result exactly (same units, same math).  `ops` is replaced by the test-only CPU stand-ins
of tests/ops_emulation.py inside every worker (there is no GPU here); the GPU path of the same code runs under NCCL."""
import ctypes
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Patch:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, T, sub, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from propainter_b200 import synth
    from propainter_b200.dist import ShardedProPainter
    from propainter_b200.inference_propainter import InferenceConfig, ProPainterPipeline
    from tests import ops_emulation
    torch.set_num_threads(2)
    ops_emulation.install(_Patch(), ctypes.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")))
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    u8, fm, md = synth.make_clip(T, 128, 128, mask="ellipse", seed=0)
    pipe = ProPainterPipeline(device="cpu")                       # same seeds on every rank -> identical weights
    # raft_clip_frames=3: every rank's pair range goes through several RAFT chunks (the correlation-volume memory cap of
    # compute_flows applies inside a shard as well)
    cfg = InferenceConfig(raft_iter=1, subvideo_length=sub, raft_clip_frames=3 if world == 2 else None)
    sp = ShardedProPainter(pipe)
    sharded = sp(torch.from_numpy(u8), fm, md, cfg, gather=True)
    part, ids = sp(torch.from_numpy(u8), fm, md, cfg)            # the sharded result: this rank's final frames only
    assert torch.equal(part, sharded[ids]) and sp.last_bytes.get('encoder_features', 0) >= 0
    if rank == 0:
        single = pipe(torch.from_numpy(u8), fm, md, cfg)
        np.savez(out_path, sharded=sharded.numpy(), single=single.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_partition_helpers():
    from propainter_b200.dist import final_frame_owner, split_range, window_owner
    from propainter_b200.inference_propainter import InferenceConfig, window_plan
    assert split_range(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)] and split_range(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    plan = window_plan(80, InferenceConfig())
    for world in (1, 2, 4, 8):
        own = window_owner(len(plan), world)
        assert own == sorted(own) and set(own) == set(range(world)) and len(own) == 16
        fin = final_frame_owner(plan, own)
        assert sorted(fin) == list(range(80)) and fin[0] == 0 and fin[79] == world - 1


@pytest.mark.parametrize("world,T,sub", [(2, 13, 6), (3, 17, 80), (4, 25, 6)])
def test_sharded_pipeline_matches_single_process(tmp_path, hostsim, world, T, sub):
    """(2 ranks, T=13, subvideo_length=6): several units in every stage and a seam between the ranks; (3 ranks, T=17, one
    sub-video): the two flow directions of the single completion unit run on different ranks, the middle rank both receives
    and forwards seam frames; (4 ranks, T=25, 5 sub-videos): ten completion tasks spread over the ranks, halos from
    non-adjacent owners (the 8-rank case T=41 was run the same way once: bit-equal).  Point-to-point exchanges only; the result must equal the single-process run bit for bit."""
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(world, _free_port(), T, sub, out), nprocs=world, join=True)
    r = np.load(out)
    assert r["sharded"].shape == (T, 128, 128, 3)
    assert np.array_equal(r["sharded"], r["single"])


def test_shard_plan_covers_everything():
    """ShardPlan invariants for the benchmarked shapes (C2: 80 frames, C4: 300 frames) at 1/2/4/8 ranks."""
    from propainter_b200.dist import ShardPlan
    from propainter_b200.inference_propainter import InferenceConfig
    for T in (80, 300, 13):
        for world in (1, 2, 4, 8):
            sp = ShardPlan(T, world, InferenceConfig())
            assert sorted(set(sp.fown)) == list(range(min(world, T))) and sp.fown == sorted(sp.fown)
            for d in (0, 1):
                assert all(o is not None for o in sp.pred_owner[d])
            assert all(o is not None for o in sp.upd_owner) and sorted(sp.final_owner) == list(range(T))
            assert sp.win_owner == sorted(sp.win_owner)
            need = sp.needs_enc()
            # a rank's windows use its own frames, <= 5 neighbour frames per side and the strided reference frames only
            for r in range(world):
                own = [i for i in range(T) if sp.fown[i] == r]
                if own and T == 300:
                    ext = [i for i in need[r] if i < own[0] - 5 or i > own[-1] + 5]
                    assert all(i % 5 == 0 for i in ext) and len(ext) <= 16, (world, r, ext)   # refs: mid +- 10k, mid a multiple of 5
            if T == 300 and world == 8:                                   # 4 sub-videos x 2 directions = one task per rank
                assert sorted(sp.s2_owner) == list(range(8))
