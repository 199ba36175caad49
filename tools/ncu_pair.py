"""One launch of each epilogue of the CTA-pair conv kernel on the dominant CIFAR shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyedm_b200 import ops
from tinyedm_b200.ops import EPI_AXPBY, EPI_MODSILU, EPI_MODSILU_BWD, EPI_SILU_BWD
dev = torch.device("cuda:0"); ops.ensure_device(dev); BF = torch.bfloat16
B, H, W, C = int(os.environ.get("NCU_B", 256)), 32, 32, 256
x = torch.randn(B, H, W, C, device=dev).to(BF)
wq = (torch.randn(C, 9 * C, device=dev) / (9 * C) ** 0.5).to(BF)
res = torch.randn(B, H, W, C, device=dev).to(BF)
mod = (torch.randn(B, C, device=dev) * 0.3 + 1).contiguous()
raw = torch.empty(B, H, W, C, device=dev, dtype=BF); out = torch.empty_like(raw)
dm = torch.zeros(B, C, device=dev); nrm = (torch.rand(B, H, W, device=dev) + 0.5).contiguous()
bn = int(os.environ.get("NCU_BN", 0))
for _ in range(2):
    ops.conv2d(x, wq, 3, C, out=out, block_n=bn)
    ops.conv2d(x, wq, 3, C, out=out, epi=EPI_AXPBY, alpha=0.4, beta=0.9, res=res, block_n=bn)
    ops.conv2d(x, wq, 3, C, out=out, epi=EPI_MODSILU, mod=mod, drop_p=0.13, seed=5, raw=raw, block_n=bn)
    ops.conv2d(x, wq, 3, C, out=out, epi=EPI_MODSILU_BWD, aux=res, mod=mod, drop_p=0.13, seed=5, d_mod=dm, block_n=bn)
    ops.conv2d(x, wq, 3, C, out=out, epi=EPI_SILU_BWD, aux=res, res=raw, beta=0.6, block_n=bn)
    ops.conv2d(x, wq, 3, C, out=out, epi=EPI_SILU_BWD, aux=res, res=raw, beta=0.6, nrm=nrm, block_n=bn if bn == 0 else 256)
torch.cuda.synchronize()
print("done")
# weight gradient of the same shape (CTA-pair kernel)
dw = torch.zeros(C, 9, C, device=dev)
for _ in range(2):
    ops.conv2d_wgrad(res, x, dw, 3)
torch.cuda.synchronize()
