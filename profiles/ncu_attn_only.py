import os, sys, torch
sys.path.insert(0, os.getcwd())
from propainter_b200 import ops
from propainter_b200.window_index import window_key_table
dev='cuda'; torch.manual_seed(0)
t, H2, W2, C = 18, 20, 36, 512
qkv = torch.randn(t, H2*W2, 3*C, device=dev); pool = torch.randn(t, 45, 2*C, device=dev)
ktab = torch.from_numpy(window_key_table(H2, W2)).to(dev)
flags = torch.zeros(16, dtype=torch.int32, device=dev); flags[[5,6,9,10,11]] = 1
for _ in range(2):
    ops.sparse_window_attn(qkv, pool, ktab, flags, t, H2*W2, 0, 2, impl="umma")
torch.cuda.synchronize()
