"""Full-size (BASELINE.json configs[1]) property test of the shipping pipeline; collected after the op / module parity
files.  Size-independent properties replace the oracle here (a CPU oracle run of the 80-frame clip takes ~10 minutes)."""
import numpy as np
import pytest
import torch

from oracle import pipeline_ref

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_full_size_properties():
    """BASELINE.json configs[1] at full size (80 x 240 x 432, shipping defaults) through properties that need no
    oracle run: (i) outside the dilated mask the composite is the input, bit-exact (inference_propainter.py:437-444);
    (ii) an all-zero mask returns the input video and leaves the RAFT flows untouched by completion (combine_flow,
    recurrent_flow_completion.py:340-347); (iii) replaying the captured graphs is bit-deterministic; (iv) the filled
    region changes when the mask moves (the fill is actually computed); (v) image propagation never grows the hole and
    only touches masked pixels (propainter.py:155-161)."""
    from propainter_b200 import synth
    from propainter_b200.inference_propainter import InferenceConfig, ProPainterPipeline
    T, H, W = 80, 240, 432
    u8, fm, md = synth.make_clip(T, H, W, mask="ellipse", seed=0)
    pipe = ProPainterPipeline(device=DEV)
    cfg = InferenceConfig()
    x = torch.from_numpy(u8)
    comp, st = pipe(x, fm, md, cfg, return_stages=True)
    a = comp.cpu().numpy()
    hole = md[0, :, 0].bool().numpy()
    assert a.shape == u8.shape and a.dtype == np.uint8
    assert np.array_equal(a[~hole], u8[~hole])                                           # (i)
    assert (a[hole] != u8[hole]).mean() > 0.5                                            # (iv) the hole was re-synthesised
    assert np.array_equal(pipe(x, fm, md, cfg).cpu().numpy(), a)                         # (iii)
    um = st["updated_masks"][0, :, 0].cpu().numpy() > 0.5
    assert not (um & ~hole).any() and um.sum() <= hole.sum()                             # (v) the hole never grows
    frames = pipeline_ref.to_float_frames(u8)[0]
    uf = st["updated_frames"][0].cpu()
    keep = ~torch.from_numpy(hole)[:, None].expand(-1, 3, -1, -1)
    assert torch.equal(uf[keep], frames[keep])
    zero = torch.zeros_like(md)
    comp0, st0 = pipe(x, zero, zero, cfg, return_stages=True)
    assert np.array_equal(comp0.cpu().numpy(), u8)                                       # (ii)
    for k in (0, 1):                                                                     # 0*pred + 1*flow == flow wherever pred is finite
        fin = torch.isfinite(st0["pred_flows"][k])
        assert torch.equal(st0["pred_flows"][k][fin], st0["gt_flows"][k][fin])
