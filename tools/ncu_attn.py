"""One launch of the attention forward and backward at the CIFAR 16x16 shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyedm_b200 import ops
dev = torch.device("cuda:0"); ops.ensure_device(dev)
B, H, heads, hd = 256, 16, 4, 64
C = heads * hd
qkv = torch.randn(B, H, H, 3 * C, device=dev).to(torch.bfloat16)
y, lse = ops.attention_forward(qkv, heads, need_lse=True)
g = torch.randn_like(y)
for _ in range(2):
    ops.attention_forward(qkv, heads, need_lse=True)
    ops.attention_backward(qkv, y, g, lse, heads)
torch.cuda.synchronize()
