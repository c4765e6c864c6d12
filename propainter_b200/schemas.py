"""state_dict schemas of the three nets on the hot path (key names / shapes == reference).

RAFT:  RAFT/raft.py:48-56, RAFT/extractor.py:6-58,118-165, RAFT/update.py:6-14,33-43,79-97,114-125
RFC :  model/recurrent_flow_completion.py:9-28,46-65,148-160,172-190,203-264
GEN :  model/propainter.py:34-54,72-101,193-216,235-304; model/modules/sparse_transformer.py:7-17,
       34-47,64-72,117-153,284-292,321-329
The golden manifest ``tests/golden/state_dict_manifest.json`` (dumped from the reference modules in
the authoring container) pins these in the CPU test-suite.
"""
import torch

from ._params import Schema


def raft_schema():
    S = Schema()
    for enc, bn in (("fnet", False), ("cnet", True)):
        def norm(p, c):
            if bn:
                S.batchnorm(p, c)          # InstanceNorm2d (fnet) has no parameters / buffers
        S.conv(f"{enc}.conv1", 3, 64, 7, gain=1.4)
        norm(f"{enc}.norm1", 64)
        cin = 64
        for li, dim, stride in ((1, 64, 1), (2, 96, 2), (3, 128, 2)):
            for bi in (0, 1):
                p = f"{enc}.layer{li}.{bi}"
                S.conv(p + ".conv1", cin, dim, 3, gain=1.4)
                S.conv(p + ".conv2", dim, dim, 3, gain=1.4)
                norm(p + ".norm1", dim)
                norm(p + ".norm2", dim)
                if bi == 0 and stride != 1:
                    norm(p + ".norm3", dim)
                    S.conv(p + ".downsample.0", cin, dim, 1)
                    if bn:
                        S.alias(p + ".downsample.1", p + ".norm3")   # same module object upstream
                cin = dim
        S.conv(f"{enc}.conv2", 128, 256, 1)
    u = "update_block."
    S.conv(u + "encoder.convc1", 324, 256, 1, gain=1.4)
    S.conv(u + "encoder.convc2", 256, 192, 3, gain=1.4)
    S.conv(u + "encoder.convf1", 2, 128, 7, gain=1.4)
    S.conv(u + "encoder.convf2", 128, 64, 3, gain=1.4)
    S.conv(u + "encoder.conv", 256, 126, 3, gain=1.4)
    for tag, k in (("1", (1, 5)), ("2", (5, 1))):
        for gate in "zrq":
            S.conv(f"{u}gru.conv{gate}{tag}", 384, 128, k)
    S.conv(u + "flow_head.conv1", 128, 256, 3, gain=1.4)
    S.conv(u + "flow_head.conv2", 256, 2, 3, gain=0.05)   # keeps random-init flow within a few px
    S.conv(u + "mask.0", 128, 256, 3, gain=1.4)
    S.conv(u + "mask.2", 256, 576, 1)
    return S


def _offset_net(S, p, cond_ch, ch=128, groups=16):
    S.conv(p + ".conv_offset.0", cond_ch, ch, 3, gain=1.3)
    S.conv(p + ".conv_offset.2", ch, ch, 3, gain=1.3)
    S.conv(p + ".conv_offset.4", ch, ch, 3, gain=1.3)
    # the reference zero-initialises this layer (offsets == 0, modulation == 0.5), which never
    # exercises the deformable gather; our synthetic init keeps it live (SURVEY.md §7 "hard parts").
    S.conv(p + ".conv_offset.6", ch, 27 * groups, 3, gain=0.7)


def rfc_schema():
    S = Schema()
    S.conv("downsample.0", 3, 32, (1, 5, 5), gain=1.3)
    for enc, specs in (("encoder1", ((0, 32, 32), (2, 32, 64))), ("encoder2", ((0, 64, 64), (2, 64, 128)))):
        for i, cin, cout in specs:
            S.conv(f"{enc}.{i}.conv1.0", cin, cout, (1, 3, 3), gain=1.3)
            S.conv(f"{enc}.{i}.conv2.0", cout, cout, (3, 1, 1), gain=1.3)
    for i in (0, 2, 4):
        S.conv(f"mid_dilation.{i}", 128, 128, (1, 3, 3), gain=1.3)
    fp = "feat_prop_module."
    for i, name in enumerate(("backward_", "forward_")):
        S.conv(fp + "deform_align." + name, 256, 128, 3)
        _offset_net(S, fp + "deform_align." + name, 384)
        S.conv(fp + f"backbone.{name}.0", (2 + i) * 128, 128, 3, gain=1.3)
        S.conv(fp + f"backbone.{name}.2", 128, 128, 3, gain=0.5)
    S.conv(fp + "fusion", 256, 128, 1)
    S.conv("decoder2.0", 128, 128, 3, gain=1.3)
    S.conv("decoder2.2.conv", 128, 64, 3, gain=1.3)
    S.conv("decoder1.0", 64, 64, 3, gain=1.3)
    S.conv("decoder1.2.conv", 64, 32, 3, gain=1.3)
    S.conv("upsample.0", 32, 32, 3, gain=1.3)
    S.conv("upsample.2.conv", 32, 2, 3)
    # edge head: training-only upstream (:301-305) but part of the strict state_dict
    S.conv("edgeDetector.projection.0", 2, 16, 3)
    S.conv("edgeDetector.mid_layer_1.0", 16, 16, 3)
    S.conv("edgeDetector.mid_layer_2.0", 16, 16, 3)
    S.conv("edgeDetector.out_layer", 16, 1, 1)
    return S


def generator_schema(depths=8, hidden=512, channel=128):
    S = Schema()
    enc = ((0, 5, 64, 1), (2, 64, 64, 1), (4, 64, 128, 1), (6, 128, 256, 1), (8, 256, 384, 1),
           (10, 640, 512, 2), (12, 768, 384, 4), (14, 640, 256, 8), (16, 512, 128, 1))
    for i, cin, cout, g in enc:
        S.conv(f"encoder.layers.{i}", cin, cout, 3, groups=g, gain=1.3)
    S.conv("decoder.0.conv", channel, 128, 3, gain=1.3)
    S.conv("decoder.2", 128, 64, 3, gain=1.3)
    S.conv("decoder.4.conv", 64, 64, 3, gain=1.3)
    S.conv("decoder.6", 64, 3, 3, gain=0.7)
    S.linear("ss.embedding", 49 * channel, hidden)
    S.linear("sc.embedding", hidden, 49 * channel, gain=1.0)
    S.conv("sc.bias_conv", channel, channel, 3, gain=0.6)
    fp = "feat_prop_module."
    for name in ("backward_1", "forward_1"):
        S.conv(fp + "deform_align." + name, channel, channel, 3)
        _offset_net(S, fp + "deform_align." + name, 2 * channel + 5)
    for name in ("backward_1", "forward_1"):
        S.conv(fp + f"backbone.{name}.0", 2 * channel + 2, channel, 3, gain=1.3)
        S.conv(fp + f"backbone.{name}.2", channel, channel, 3, gain=0.5)
    S.conv(fp + "fuse.0", 2 * channel + 2, channel, 3, gain=1.3)
    S.conv(fp + "fuse.2", channel, channel, 3, gain=0.5)
    from .window_index import rolled_valid_index
    for i in range(depths):
        p = f"transformers.transformer.{i}."
        S.add(p + "attention.valid_ind_rolled", (148,), kind="buffer", dtype=torch.int64,
              init=("const", rolled_valid_index((5, 9))))
        for n in ("key", "query", "value", "proj"):
            S.linear(p + "attention." + n, hidden, hidden, gain=1.0 if n != "proj" else 0.5)
        S.add(p + "attention.pool_layer.weight", (hidden, 1, 4, 4), init=("normal_mean", 1.0 / 16, 0.02))
        S.add(p + "attention.pool_layer.bias", (hidden,), init=("normal", 0.02))
        S.affine(p + "norm1", hidden)
        S.affine(p + "norm2", hidden)
        S.linear(p + "mlp.fc1.0", hidden, 1960)
        S.linear(p + "mlp.fc2.1", 1960, hidden, gain=0.5)
    return S
