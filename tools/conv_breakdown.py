"""Per-shape time of every conv / wgrad launch in one CIFAR training step (CUDA events around each C-ABI call)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tinyedm_b200 as T
from tinyedm_b200 import ops
import tinyedm_b200.engine as engine_mod
from tinyedm_b200.configs import CIFAR10, build_edm
dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(os.environ.get("B", "256"))
model = build_edm(CIFAR10).to(dev).train()
with torch.no_grad(): model.denoiser.gain_out.fill_(1.0)
opt = model.configure_optimizers()["optimizer"]
x = (0.5 * torch.randn(B, 3, 32, 32, device=dev)).clamp(-1, 1); y = torch.zeros(B, dtype=torch.long, device=dev)
def step():
    opt.zero_grad(set_to_none=True); model.training_step((x, y), 0).backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
events = []
oc, ow = ops.conv2d, ops.conv2d_wgrad
def conv(xx, w, k, cout, **kw):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = oc(xx, w, k, cout, **kw); b.record()
    Bn, H, W, cin = xx.shape
    events.append((("conv", H, cin, cout, k, kw.get("epi", 0), kw.get("nrm") is not None), 2.0 * Bn * H * W * cin * cout * k * k, a, b))
    return out
def wgrad(g, xx, dw, k, **kw):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ow(g, xx, dw, k, **kw); b.record()
    Bn, H, W, cout = g.shape; cin = xx.shape[3]
    events.append((("wgrad", H, cin, cout, k, 0, False), 2.0 * Bn * H * W * cin * cout * k * k, a, b))
engine_mod.ops.conv2d = conv; engine_mod.ops.conv2d_wgrad = wgrad
for _ in range(3): step()
torch.cuda.synchronize()
engine_mod.ops.conv2d = oc; engine_mod.ops.conv2d_wgrad = ow
agg = collections.OrderedDict()
for key, fl, a, b in events:
    n, t, f = agg.get(key, (0, 0.0, 0.0))
    agg[key] = (n + 1, t + a.elapsed_time(b), f + fl)
tot_t = sum(v[1] for v in agg.values()) / 3; tot_f = sum(v[2] for v in agg.values()) / 3
print(f"all conv launches: {tot_t:.2f} ms/step, {tot_f/1e12:.2f} TFLOP/step -> {tot_f/tot_t/1e9:.0f} TF/s")
names = {0: "plain", 1: "modsilu", 2: "axpby", 3: "modsilu'", 4: "silu'"}
for key, (n, t, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    kind, H, cin, cout, k, epi, pn = key
    ideal = f / 3 / 1.4e15 * 1e3
    print(f"{kind:5s} {H:2d}x{H:<2d} {cin:4d}->{cout:<4d} k{k} {names[epi]+('+pn' if pn else ''):11s} n/step={n//3:3d} {t/3:7.3f} ms/step  {f/t/1e9:6.0f} TF/s  avg {t/n*1e3:6.1f} us  (at 1400 TF/s: {ideal:.3f} ms, excess {t/3-ideal:.3f})")
