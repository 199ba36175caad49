"""Kernel-level breakdown of one training step of the secondary configs (MNIST, ImageNet-512 latents) and their
CUDA-graph step times. Run under gpurun; writes gpurun_out/profile_configs.txt."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import tinyedm_b200 as T
from tinyedm_b200.configs import IMAGENET, MNIST, build_edm
dev = torch.device("cuda:0")
os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/profile_configs.txt", "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); log.write(s + "\n"); log.flush()
for name, cfg, B, gflop in (("mnist", MNIST, 128, 20.11), ("imagenet_latent", IMAGENET, int(os.environ.get("IN_B", "64")), 192.9)):
    torch.manual_seed(0)
    model = build_edm(cfg).to(dev).train()
    with torch.no_grad(): model.denoiser.gain_out.fill_(1.0)
    opt = model.configure_optimizers()["optimizer"]
    for g in opt.param_groups: g["lr"] = 1e-5
    C, H, W = cfg["image"]
    x = (0.5 * torch.randn(B, C, H, W, device=dev)).clamp(-1, 1)
    y = torch.randint(0, cfg["embedding"]["num_classes"], (B,), device=dev)
    def step():
        opt.zero_grad(set_to_none=True); loss = model.training_step((x, y), 0); loss.backward(); opt.step(); return loss
    for _ in range(3): step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(); torch.cuda.synchronize()
    evs = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in evs)
    P(f"== {name} B={B}: total device time {tot/1e3:.2f} ms")
    for e in evs[:16]:
        P(f"{e.device_time_total/1e3:9.3f} ms {100*e.device_time_total/tot:5.1f}%  n={e.count:4d}  avg {e.device_time_total/e.count:8.1f} us  {e.key[:120]}")
    gstep = T.GraphedTrainStep(model, opt, (x, y))
    for _ in range(3): gstep((x, y))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 10
    for _ in range(n): gstep((x, y))
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    P(f"{name}: graphed step {dt*1e3:.2f} ms = {B/dt:.0f} img/s = {3*gflop*B/dt/1e3:.0f} TFLOP/s (graph: {gstep.graph is not None}, {gstep.error})")
    del model, opt, gstep
    torch.cuda.empty_cache()
