"""One launch of every tcgen05 attention kernel at the configs' shapes (for ncu): the kernels specialised for head_dim 64
(CIFAR 16x16 / 8x8) and the generic pair (MNIST 14x14 / 7x7, ImageNet-latent 16x16 / 8x8) with its normalisation pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyedm_b200 import ops
dev = torch.device("cuda:0"); ops.ensure_device(dev)
for (B, H, heads, hd) in [(256, 16, 4, 64), (256, 8, 4, 64)]:
    C = heads * hd
    qkv = torch.randn(B, H, H, 3 * C, device=dev).to(torch.bfloat16)
    y, lse = ops.attention_forward(qkv, heads, need_lse=True)
    g = torch.randn_like(y)
    ops.attention_backward(qkv, y, g, lse, heads)
for (B, H, heads, hd) in [(128, 14, 4, 64), (128, 7, 4, 128), (176, 16, 4, 144), (176, 8, 4, 192)]:
    C = heads * hd
    qkv = torch.randn(B, H, H, 3 * C, device=dev).to(torch.bfloat16)
    qn, norms = ops.qkv_normalize(qkv, heads)
    y, lse = ops.attention_forward_normalized(qn, heads, need_lse=True)
    g = torch.randn_like(y)
    ops.attention_backward_normalized(qn, norms, y, g, lse, heads)
torch.cuda.synchronize()
print("done")
