"""For ncu: the transposed CTA-pair weight-gradient kernel at three ImageNet-latent shapes (B = 64), then the weight_prep
kernels (norms, tiles, backward) over the CIFAR weight bank. Each kernel is launched twice; profile the second launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyedm_b200 import configs, ops
from tinyedm_b200.networks import Denoiser
dev = torch.device("cuda:0"); ops.ensure_device(dev); BF = torch.bfloat16
torch.manual_seed(0)
for (B, H, Cin, Cout, ks) in [(64, 64, 192, 192, 3), (64, 32, 384, 384, 3), (64, 16, 576, 576, 3)]:
    x = torch.randn(B, H, H, Cin, device=dev).to(BF)
    g = torch.randn(B, H, H, Cout, device=dev).to(BF)
    dw = torch.zeros(Cout, ks * ks, Cin, device=dev)
    for _ in range(2):
        ops.conv2d_wgrad(g, x, dw, ks, accumulate=True)
    torch.cuda.synchronize()
den = Denoiser(**configs.CIFAR10["denoiser"]).to(dev)
eng = den.engine; eng._ensure_device(dev); bank = eng.bank
bank.ensure_grad_buffers()
for _ in range(2):
    bank.invalidate(); bank.prepare(True)
    bank.backward()
torch.cuda.synchronize()
print("done")
