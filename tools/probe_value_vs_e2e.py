"""Why does bench.py's device-input `value` come out 1-2.5 % BELOW its host-input `e2e` on every box? Alternates the two
timed loops (same graph-replayed CIFAR step, 30 steps each) several times in one process and prints ms/step of each pass, so
an ordering effect (the first region after capture) can be told from a real difference between the two input paths.

    python tools/probe_value_vs_e2e.py > gpurun_out/value_vs_e2e.txt
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tinyedm_b200 as T
from tinyedm_b200.configs import CIFAR10, build_edm

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B, STEPS = 256, int(os.environ.get("STEPS", "30"))
torch.manual_seed(0)
model = build_edm(CIFAR10).to(dev).train()
with torch.no_grad():
    model.denoiser.gain_out.fill_(1.0)
opt = model.configure_optimizers()["optimizer"]
for g in opt.param_groups: g["lr"] = 2e-5
gen = torch.Generator().manual_seed(1)
host = [((0.5 * torch.randn(B, 3, 32, 32, generator=gen)).clamp_(-1, 1).pin_memory(), torch.zeros(B, dtype=torch.long).pin_memory()) for _ in range(4)]
devb = [(x.to(dev), y.to(dev)) for x, y in host]
step = T.GraphedTrainStep(model, opt, devb[0])
assert step.graph is not None, step.error
loss_host = torch.zeros(STEPS, pin_memory=True)
for _ in range(3): step(devb[0])
torch.cuda.synchronize()


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(STEPS): fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / STEPS


def dev_in(i): step(devb[i % 4])
def host_in(i):
    loss = step(host[i % 4])
    loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)
def dev_in_readback(i):
    loss = step(devb[i % 4])
    loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)


order = [("device", dev_in), ("host", host_in), ("device", dev_in), ("host", host_in), ("device+readback", dev_in_readback),
         ("host", host_in), ("device", dev_in), ("device", dev_in), ("host", host_in)]
t0 = time.time()
for name, fn in order:
    ms = timed(fn)
    print(f"t+{time.time() - t0:6.2f}s  {name:16s} {ms:7.3f} ms/step  {B / ms * 1e3:8.0f} img/s", flush=True)
time.sleep(3.0)                      # an idle pause, then the same again: does a cold start change the first pass?
print("-- after 3 s idle")
for name, fn in order[:4]:
    ms = timed(fn)
    print(f"t+{time.time() - t0:6.2f}s  {name:16s} {ms:7.3f} ms/step  {B / ms * 1e3:8.0f} img/s", flush=True)
