"""CPU tests of the drop-in modules' host-side plumbing (weight packing, concat-buffer channel orders, scan
bookkeeping, stage scheduling, compositing order) against the oracle, with ``propainter_b200.ops`` replaced
by the test-only CPU stand-ins of tests/ops_emulation.py.  The kernels themselves are checked on the GPU
(test_gpu_*.py); their per-element rules on the CPU in test_elem_hostsim.py."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import flowcomp_ref, generator_ref, ops_ref, pipeline_ref, raft_ref
from tests import ops_emulation


@pytest.fixture()
def emu(monkeypatch, hostsim):
    ops_emulation.install(monkeypatch, hostsim)


def rel_err(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def test_raft_plumbing(emu):
    from propainter_b200 import synth
    from propainter_b200.model.modules.flow_comp_raft import RAFT_bi
    net = RAFT_bi(None, "cpu", seed=1)
    u8, _, _ = synth.make_clip(3, 128, 144, seed=3)
    frames = pipeline_ref.to_float_frames(u8)
    fw, bw = net(frames, iters=3)
    rf, rb = raft_ref.raft_bi(net.fix_raft.state_dict(), frames, 3)
    assert rel_err(fw, rf) < 1e-4 and rel_err(bw, rb) < 1e-4, (rel_err(fw, rf), rel_err(bw, rb))
    lo, up = net.fix_raft(frames[0, :2], frames[0, 1:3], iters=2, test_mode=True)
    rlo, rup = raft_ref.raft_forward(net.fix_raft.state_dict(), frames[0, :2], frames[0, 1:3], 2, return_lowres=True)
    assert rel_err(up, rup) < 1e-4 and rel_err(lo, rlo) < 1e-4


def test_flow_completion_plumbing(emu):
    from propainter_b200.model.recurrent_flow_completion import RecurrentFlowCompleteNet
    net = RecurrentFlowCompleteNet(None, seed=2)
    gen = torch.Generator().manual_seed(0)
    T, H, W = 6, 32, 48
    flows = (torch.randn(2, T - 1, 2, H, W, generator=gen) * 3, torch.randn(2, T - 1, 2, H, W, generator=gen) * 3)
    masks = torch.zeros(2, T, 1, H, W)
    masks[..., 8:24, 12:36] = 1
    pred, _ = net.forward_bidirect_flow(flows, masks)
    ref = flowcomp_ref.forward_bidirect_flow(net.state_dict(), flows, masks)
    assert rel_err(pred[0], ref[0]) < 1e-4 and rel_err(pred[1], ref[1]) < 1e-4, (rel_err(pred[0], ref[0]), rel_err(pred[1], ref[1]))


def test_flow_completion_both_conv_plans(emu, monkeypatch):
    """config.UMMA_CONV on (segmented inputs, hoisted frame-only conv shares, gather + 1x1 GEMM) and off (cuDNN convs +
    concat buffers + pp_deform_align) are two execution plans of the same scan: both must match the oracle."""
    from propainter_b200 import config
    from propainter_b200.model.recurrent_flow_completion import RecurrentFlowCompleteNet
    net = RecurrentFlowCompleteNet(None, seed=2)
    gen = torch.Generator().manual_seed(0)
    flows = (torch.randn(1, 5, 2, 32, 48, generator=gen), torch.randn(1, 5, 2, 32, 48, generator=gen))
    masks = torch.zeros(1, 6, 1, 32, 48)
    masks[..., 8:24, 12:36] = 1
    ref = flowcomp_ref.forward_bidirect_flow(net.state_dict(), flows, masks)
    from propainter_b200 import autotune
    # hybrid: library convs + gather / 1x1-GEMM deformable conv; hoisted: + frame-only conv shares once per scan; 3 / 4: the
    # autotune candidates "hoisted" and "hoisted + gather/GEMM"
    for flag in (True, False, "hybrid", "hoisted", 3, 4):
        if isinstance(flag, int) and not isinstance(flag, bool):
            monkeypatch.setattr(config, "UMMA_CONV", "auto")
            monkeypatch.setattr(autotune, "pick", lambda key, variants, *a, _i=flag, **k: variants[_i if key[0] == "rfc_prop" else 0](*a))
        else:
            monkeypatch.setattr(config, "UMMA_CONV", flag)
        out, _ = net.forward_bidirect_flow(flows, masks)
        for k in (0, 1):
            assert rel_err(out[k], ref[k]) < 1e-4, (flag, k, rel_err(out[k], ref[k]))


@pytest.mark.parametrize("H,W,t,lt,alt", [(64, 96, 4, 3, False), (128, 128, 3, 2, False), (64, 64, 2, 1, False), (64, 96, 4, 3, True),
                                          (64, 96, 4, 3, "cudnn"), (64, 64, 2, 1, "cudnn"), (64, 96, 4, 3, "hybrid"),
                                          (64, 96, 4, 3, "hoisted"), (64, 64, 2, 1, "hoisted")])
def test_generator_plumbing(emu, monkeypatch, H, W, t, lt, alt):
    from propainter_b200 import autotune, config
    from propainter_b200.model.propainter import InpaintGenerator
    if alt == "cudnn":                                  # the library-conv plan of the propagation scan (config.UMMA_CONV off)
        monkeypatch.setattr(config, "UMMA_CONV", False)
        alt = False
    if alt in ("hybrid", "hoisted"):                    # library convs + pp_deform_gather / 1x1 pp_conv2d_umma deformable conv
        monkeypatch.setattr(config, "UMMA_CONV", alt)   # (hoisted: + the frame-only conv shares once per scan)
        alt = False
    if alt:      # force the alternative execution plans autotune may pick on the GPU (per-group dense encoder convs)
        monkeypatch.setattr(autotune, "pick", lambda key, variants, *a, **k: variants[0 if key[0] == "conv_relu" else -1](*a))
    net = InpaintGenerator(seed=3)
    gen = torch.Generator().manual_seed(1)
    frames = torch.rand(1, t, 3, H, W, generator=gen) * 2 - 1
    sm = lambda z: F.avg_pool2d(z.view(-1, 2, H, W), 9, 1, 4).view(z.shape)
    flows = (sm(torch.randn(1, lt - 1, 2, H, W, generator=gen) * 12), sm(torch.randn(1, lt - 1, 2, H, W, generator=gen) * 12))
    masks = torch.zeros(1, t, 1, H, W)
    masks[..., H // 4:H // 2, W // 3:2 * W // 3] = 1
    upd = masks * (torch.rand(1, t, 1, H, W, generator=gen) > 0.5).float()
    mf = frames * (1 - masks)
    out = net(mf, flows, masks, upd, lt)
    ref = generator_ref.generator_forward(net.state_dict(), mf, flows, masks, upd, lt)
    assert out.shape == ref.shape and rel_err(out, ref) < 2e-4, rel_err(out, ref)


def test_pipeline_plumbing(emu):
    from propainter_b200 import synth
    from propainter_b200.inference_propainter import InferenceConfig, ProPainterPipeline
    T, H, W = 13, 128, 128
    u8, fm, md = synth.make_clip(T, H, W, mask="ellipse", seed=0)
    pipe = ProPainterPipeline(device="cpu")
    cfg = InferenceConfig(raft_iter=1, subvideo_length=6)              # forces halo chunking in stages 2/3
    comp, st = pipe(torch.from_numpy(u8), fm, md, cfg, return_stages=True)
    ref, rst = pipeline_ref.run_pipeline(pipe.state_dicts(), u8, fm, md, raft_iter=1, subvideo_length=6, return_stages=True)
    for k in (0, 1):
        assert rel_err(st["gt_flows"][k], rst["gt_flows"][k]) < 1e-4
        assert rel_err(st["pred_flows"][k], rst["pred_flows"][k]) < 1e-3
    assert (st["updated_masks"] != rst["updated_masks"]).float().mean() < 2e-3
    assert ops_ref.psnr_u8(comp.numpy(), ref) > 50.0


def test_half_storage_is_accepted(emu):
    """`.half()` nets with fp16 tensors (reference --fp16, inference_propainter.py:268-270, :323-330): same call surface,
    results come back in fp16 and agree with the fp32 run to fp16 rounding (the math itself stays fp32)."""
    from propainter_b200.model.recurrent_flow_completion import RecurrentFlowCompleteNet
    torch.manual_seed(0)
    net = RecurrentFlowCompleteNet(None, seed=3)
    flows = (torch.randn(1, 3, 2, 32, 32), torch.randn(1, 3, 2, 32, 32))
    masks = (torch.rand(1, 4, 1, 32, 32) > 0.6).float()
    ref, _ = net.forward_bidirect_flow(flows, masks)
    sd32 = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.half()
    assert all(v.dtype == torch.float16 for k, v in net.state_dict().items() if v.is_floating_point())
    out, _ = net.forward_bidirect_flow((flows[0].half(), flows[1].half()), masks.half())
    assert out[0].dtype == torch.float16 and out[1].dtype == torch.float16
    for a, b in zip(out, ref):
        assert (a.float() - b).abs().max() < 2e-2 * (1 + b.abs().max())
    comb = net.combine_flow((flows[0].half(), flows[1].half()), out, masks.half())
    assert comb[0].dtype == torch.float16
    net.float().load_state_dict(sd32, strict=True)


@pytest.mark.parametrize("T,kw", [(2, {}), (3, {}), (5, dict(neighbor_length=20, ref_stride=20)), (7, dict(neighbor_length=2, ref_stride=3))])
def test_pipeline_edge_lengths(emu, T, kw):
    """Shortest clips (one flow), window / reference strides larger than the clip, and a stride-3 reference schedule:
    the driver's window plan (inference_propainter.py:159-173, :417-452) must keep matching the oracle's."""
    from propainter_b200 import synth
    from propainter_b200.inference_propainter import InferenceConfig, ProPainterPipeline
    u8, fm, md = synth.make_clip(T, 128, 128, mask="ellipse", seed=0)
    pipe = ProPainterPipeline(device="cpu")
    comp = pipe(torch.from_numpy(u8), fm, md, InferenceConfig(raft_iter=1, **kw))
    ref = pipeline_ref.run_pipeline(pipe.state_dicts(), u8, fm, md, raft_iter=1, **kw)
    assert comp.shape == (T, 128, 128, 3) and ops_ref.psnr_u8(comp.numpy(), ref) > 60.0
