"""GPU (for ncu): the dominant conv shape (B=256, 32x32, 256->256, 3x3) once per epilogue flavour."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyedm_b200 import ops
dev = torch.device("cuda:0"); ops.ensure_device(dev)
B, H, C = int(os.environ.get("B", 256)), 32, 256
torch.manual_seed(0)
x = torch.randn(B, H, H, C, device=dev).to(torch.bfloat16)
w = (torch.randn(C, 9 * C, device=dev) / (9 * C) ** 0.5).to(torch.bfloat16)
aux = torch.randn(B, H, H, C, device=dev).to(torch.bfloat16)
res = torch.randn(B, H, H, C, device=dev).to(torch.bfloat16)
mod = (torch.randn(B, C, device=dev) * 0.3 + 1).contiguous()
d_mod = torch.zeros(B, C, device=dev)
nrm = torch.rand(B * H * H, device=dev) + 0.5
raw = torch.empty_like(x); out = torch.empty_like(x)
seed = torch.zeros(1, dtype=torch.int64, device=dev)
reps = int(os.environ.get("REPS", 2))
import time
def run(name, **kw):
    for _ in range(reps):
        ops.conv2d(x, w, 3, C, out=out, **kw)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.conv2d(x, w, 3, C, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1)/reps*1e3:8.1f} us")
run("plain")
run("modsilu+drop (fwd conv1)", epi=ops.EPI_MODSILU, mod=mod, raw=raw, drop_p=0.13, seed=5, seed_ptr=seed)
run("axpby (fwd conv2)", epi=ops.EPI_AXPBY, alpha=0.4, beta=0.9, res=res)
run("modsilu_bwd (dgrad conv2)", epi=ops.EPI_MODSILU_BWD, alpha=0.4, aux=aux, mod=mod, d_mod=d_mod, drop_p=0.13, seed=5, seed_ptr=seed)
run("silu_bwd (dgrad conv1)", epi=ops.EPI_SILU_BWD, aux=aux, res=res, beta=0.9)
run("silu_bwd+pixelnorm", epi=ops.EPI_SILU_BWD, aux=aux, res=res, beta=0.9, nrm=nrm)
run("silu_bwd+pixelnorm+acc", epi=ops.EPI_SILU_BWD, aux=aux, res=res, beta=0.9, nrm=nrm, accumulate_out=True)
