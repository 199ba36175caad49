"""The 'practical bar' of SURVEY.md §8d: the reference's algorithm run EAGERLY on the same B200 (torch 2.11, bf16
autocast, cuDNN/cuBLAS kernels) — here through the oracle port (plain torch ops, same op sequence as
src/tinyedm/networks.py; the reference package itself cannot be imported on the GPU box). Training step img/s at
B = 256 and 32-step Heun img/s at B = 128 for the CIFAR config. Writes gpurun_out/ref_gpu_eager.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import edm2_oracle as O
dev = torch.device("cuda:0")
_mp_add = O.mp_add
O.mp_add = lambda a, b, t=0.5: _mp_add(a, b.to(a.dtype), t)   # autocast leaves mixed dtypes where the reference's islands cast
torch.backends.cudnn.benchmark = True
cfg = O.CIFAR10
gen = torch.Generator().manual_seed(42)
dp = {k: v.to(dev) for k, v in O.init_denoiser_params(cfg["denoiser"], gen, gain_out=1.0).items()}
ep = {k: v.to(dev) for k, v in O.init_embedding_params(cfg["embedding"], gen).items()}
params = [v.requires_grad_(True) for v in dp.values()] + [ep["sigma_embed.weight"].requires_grad_(True)]
weights = [v for k, v in list(dp.items()) + list(ep.items()) if k.endswith("weight")]
opt = torch.optim.Adam(params, lr=2e-5, fused=True)
B = 256
clean = (0.5 * torch.randn(B, 3, 32, 32, device=dev)).clamp(-1, 1)
def train_step():
    with torch.no_grad():
        for w in weights: O.forced_weight_norm_(w)
    noisy, sigma = O.diffuse(clean, torch.randn(B, device=dev), torch.randn_like(clean), -1.2, 1.2)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        _, e = O.embedding_forward(ep, cfg["embedding"], sigma)
        D = O.denoiser_forward(dp, cfg["denoiser"], noisy, sigma, e, drop=lambda t: torch.nn.functional.dropout(t, 0.13, True))
    loss = O.training_loss(O.loss_weight(sigma, 0.5), D.float(), clean)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
def timeit(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
t_train = timeit(train_step, 10)
Bs = 128
x = torch.randn(Bs, 3, 32, 32, device=dev); sig = torch.full((Bs,), 1.5, device=dev)
@torch.no_grad()
def nfe():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        _, e = O.embedding_forward(ep, cfg["embedding"], sig)
        return O.denoiser_forward(dp, cfg["denoiser"], x, sig, e)
t_nfe = timeit(nfe, 10)
out = {"what": "oracle port (reference op sequence) eager on B200, bf16 autocast, cudnn.benchmark", "train_img_s": B / t_train,
       "train_ms_per_step": t_train * 1e3, "nfe_ms_b128": t_nfe * 1e3, "heun32_img_s": Bs / (63 * t_nfe)}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ref_gpu_eager.json", "w"))
