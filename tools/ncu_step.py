"""Two CIFAR training steps (B=256) for ncu captures filtered by kernel name (-k regex:...)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyedm_b200.configs import CIFAR10, build_edm
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = build_edm(CIFAR10).to(dev).train()
with torch.no_grad():
    model.denoiser.gain_out.fill_(1.0)
opt = model.configure_optimizers()["optimizer"]
for g in opt.param_groups: g["lr"] = 2e-5
B = int(os.environ.get("B", "256"))
x = (0.5 * torch.randn(B, 3, 32, 32, device=dev)).clamp(-1, 1)
y = torch.zeros(B, dtype=torch.long, device=dev)
for _ in range(int(os.environ.get("STEPS", "2"))):
    opt.zero_grad(set_to_none=True)
    loss = model.training_step((x, y), 0)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("done", float(loss))
