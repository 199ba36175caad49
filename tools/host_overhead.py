"""GPU: host-side cost of one training step / one NFE, measured at a tiny batch where device time is negligible, plus
per-shape timings of the conv forward / wgrad kernels of the CIFAR net at B=256."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tinyedm_b200 as T
from tinyedm_b200 import ops
from tinyedm_b200.configs import CIFAR10, build_edm
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = build_edm(CIFAR10).to(dev).train()
with torch.no_grad(): model.denoiser.gain_out.fill_(1.0)
opt = model.configure_optimizers()["optimizer"]
for g in opt.param_groups: g["lr"] = 2e-5
for B in (4, 256):
    x = (0.5 * torch.randn(B, 3, 32, 32, device=dev)).clamp(-1, 1); y = torch.zeros(B, dtype=torch.long, device=dev)
    def step():
        opt.zero_grad(set_to_none=True)
        loss = model.training_step((x, y), 0); loss.backward(); opt.step()
    for _ in range(3): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): step()
    t_host = (time.perf_counter() - t0) / 5
    torch.cuda.synchronize(); t_wall = (time.perf_counter() - t0) / 5
    print(f"B={B}: host {t_host*1e3:.2f} ms/step, wall {t_wall*1e3:.2f} ms/step")
    # split: forward only / backward only host time
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loss = model.training_step((x, y), 0); t1 = time.perf_counter()
    loss.backward(); t2 = time.perf_counter(); opt.step(); t3 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"   host: fwd {1e3*(t1-t0):.2f}  bwd {1e3*(t2-t1):.2f}  opt {1e3*(t3-t2):.2f} ms")

def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("per-shape conv / wgrad at B=256 (count per train step in brackets)")
shapes = [(32, 256, 256, 3, "fwd x12, dgrad x12, wgrad x12"), (32, 512, 256, 3, "fwd/wgrad x3; dgrad is 256->512"), (32, 256, 512, 3, "dgrad of 512->256 x3"),
          (16, 256, 256, 3, "x14"), (16, 512, 256, 3, "x3"), (16, 256, 512, 3, "dgrad x3"), (8, 256, 256, 3, "x16"), (8, 512, 256, 3, "x3"), (8, 256, 512, 3, "dgrad x3"),
          (32, 512, 256, 1, "x3"), (16, 512, 256, 1, "x3"), (8, 512, 256, 1, "x3"), (16, 256, 768, 1, "qkv x5"), (16, 768, 256, 1, "qkv dgrad x5"), (16, 256, 256, 1, "out x5"),
          (8, 256, 768, 1, "qkv x6"), (8, 768, 256, 1, "x6"), (8, 256, 256, 1, "out x6")]
for (hw, cin, cout, k, note) in shapes:
    B = 256
    x = torch.randn(B, hw, hw, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, k * k * cin, device=dev) / (cin * k * k) ** 0.5).to(torch.bfloat16)
    g = torch.randn(B, hw, hw, cout, device=dev).to(torch.bfloat16)
    dw = torch.empty(cout, k * k, cin, device=dev)
    out = torch.empty(B, hw, hw, cout, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * B * hw * hw * cin * cout * k * k
    t = bench(lambda: ops.conv2d(x, w, k, cout, out=out))
    tw = bench(lambda: ops.conv2d_wgrad(g, x, dw, k)) if cin % 64 == 0 and cout % 64 == 0 else float("nan")
    print(f"  {hw:2d}x{hw:<2d} {cin:4d}->{cout:<4d} k{k}: conv {t*1e3:7.1f} us {fl/t/1e9:7.0f} TF/s | wgrad {tw*1e3:7.1f} us {fl/tw/1e9:7.0f} TF/s   [{note}]")
