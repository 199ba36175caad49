"""GPU: per-kernel time breakdown of one training step and one sampling NFE (torch.profiler/CUPTI sees every kernel of the
process, including libtinyedm_b200.so's). Writes gpurun_out/kernels_train.txt / kernels_sample.txt."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import tinyedm_b200 as T
from tinyedm_b200.configs import CIFAR10, build_edm

dev = torch.device("cuda:0")
os.makedirs("gpurun_out", exist_ok=True)
B = int(os.environ.get("B", "256"))
torch.manual_seed(0)
model = build_edm(CIFAR10).to(dev).train()
with torch.no_grad():
    model.denoiser.gain_out.fill_(1.0)
opt = model.configure_optimizers()["optimizer"]
for g in opt.param_groups: g["lr"] = 2e-5
x = (0.5 * torch.randn(B, 3, 32, 32, device=dev)).clamp(-1, 1)
y = torch.zeros(B, dtype=torch.long, device=dev)

def step():
    opt.zero_grad(set_to_none=True)
    loss = model.training_step((x, y), 0)
    loss.backward()
    opt.step()

for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter(); step(); t_cpu = time.perf_counter() - t0
torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print(f"train step: host launch time {t_cpu*1e3:.1f} ms, wall {t_all*1e3:.1f} ms")

def table(prof, path, n=45):
    evs = [e for e in prof.key_averages() if e.device_time_total > 0]
    evs.sort(key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in evs)
    with open(path, "w") as f:
        f.write(f"total device time {tot/1e3:.2f} ms\n")
        for e in evs[:n]:
            f.write(f"{e.device_time_total/1e3:9.3f} ms {100*e.device_time_total/tot:5.1f}%  n={e.count:4d}  avg {e.device_time_total/e.count:8.1f} us  {e.key[:110]}\n")
    print(open(path).read())

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
table(prof, "gpurun_out/kernels_train.txt")

# the graph-replayed step: busy time vs span (launch gaps that remain inside the graph)
def gaps(prof, tag):
    evs = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_time_total > 0 and e.time_range.end > e.time_range.start))
    if not evs:
        return
    span = evs[-1][1] - evs[0][0]
    busy, cur_end, idle_small, idle_big, n_big, prev = 0.0, evs[0][0], 0.0, 0.0, 0, ""
    for a, b, name in evs:
        if a > cur_end:
            g = a - cur_end
            if g > 5:
                idle_big += g; n_big += 1
                print(f"   gap {g:8.1f} us at +{(cur_end - evs[0][0])/1e3:7.2f} ms between [{prev[:70]}] and [{name[:70]}]")
            else: idle_small += g
        busy += max(0.0, b - max(a, cur_end))
        if b > cur_end: cur_end = b; prev = name
    print(f"{tag}: span {span/1e3:.2f} ms, busy {busy/1e3:.2f} ms, idle in gaps<=5us {idle_small/1e3:.2f} ms, idle in {n_big} gaps>5us {idle_big/1e3:.2f} ms, {len(evs)} device activities")
gaps(prof, "eager step")
gstep = T.GraphedTrainStep(model, opt, (x, y))
print("graph captured:", gstep.graph is not None, gstep.error, "launches/step", gstep.launches_per_step)
for _ in range(3): gstep((x, y))
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): gstep((x, y))
torch.cuda.synchronize(); print(f"graphed step: {(time.perf_counter()-t0)*100:.2f} ms/step wall")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gstep((x, y)); torch.cuda.synchronize()
gaps(prof, "graphed step")

# sampling: one network evaluation, eval mode, batch 128, class conditional
sm = build_edm(CIFAR10, num_classes=10, dropout_rate=0.0).to(dev).eval()
with torch.no_grad():
    sm.denoiser.gain_out.fill_(1.0)
xs = torch.randn(128, 3, 32, 32, device=dev); lab = torch.randint(0, 10, (128, 1), device=dev)
sig = torch.tensor(1.5, device=dev)
with torch.no_grad():
    for _ in range(3): sm(xs, sig, lab)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); sm(xs, sig, lab); t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    print(f"sample NFE: host launch time {t_cpu*1e3:.2f} ms, wall {t_all*1e3:.2f} ms")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        sm(xs, sig, lab); torch.cuda.synchronize()
table(prof, "gpurun_out/kernels_sample.txt", 25)
