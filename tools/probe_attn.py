"""Timing of the attention forward/backward at the CIFAR shapes (run with TEDM_ATTN_TC=0/1 to compare the kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyedm_b200 import ops
dev = torch.device("cuda:0"); ops.ensure_device(dev)
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (B, H, heads, hd) in [(256, 16, 4, 64), (256, 8, 4, 64), (128, 16, 4, 64), (128, 8, 4, 64)]:
    C = heads * hd; S = H * H
    qkv = torch.randn(B, H, H, 3 * C, device=dev).to(torch.bfloat16)
    y, lse = ops.attention_forward(qkv, heads, need_lse=True)
    g = torch.randn_like(y)
    tf = bench(lambda: ops.attention_forward(qkv, heads, need_lse=True))
    tb = bench(lambda: ops.attention_backward(qkv, y, g, lse, heads))
    fl = 4.0 * B * heads * S * S * hd
    print(f"TC={os.environ.get('TEDM_ATTN_TC','1')} B{B} S{S} hd{hd}: fwd {tf*1e3:7.1f} us {fl/tf/1e9:6.1f} TF/s | bwd {tb*1e3:7.1f} us {2.5*fl/tb/1e9:6.1f} TF/s")
