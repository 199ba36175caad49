"""GPU probe: transposed CTA-pair wgrad kernel (k slabs on M, output channels on N) vs the single-CTA one (splits=-1) and
torch, at the ImageNet-latent / MNIST channel counts. TEDM_WGRAD_N192=0 rounds N up to 128-multiples instead of N = 192."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tinyedm_b200 import ops
os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/probe_wgrad_t%s.log" % os.environ.get("TEDM_WGRAD_N192", ""), "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); log.write(s + "\n"); log.flush()
dev = torch.device("cuda:0"); ops.ensure_device(dev); BF = torch.bfloat16
torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = False
def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()
def ref_wgrad(g, x, ks):
    xr = x.float().permute(0, 3, 1, 2).contiguous()
    Cout, Cin = g.shape[3], x.shape[3]
    w = torch.zeros(Cout, Cin, ks, ks, device=dev, requires_grad=True)
    (gw,) = torch.autograd.grad(F.conv2d(xr, w, padding="same"), w, g.float().permute(0, 3, 1, 2))
    return gw.permute(0, 2, 3, 1).reshape(Cout, ks * ks, Cin)
ok_all = True
for (B, H, W, Cin, Cout, ks) in [(9, 8, 8, 64, 128, 3), (2, 28, 28, 128, 128, 3), (3, 16, 16, 128, 320, 3), (3, 64, 64, 192, 192, 3),
                                 (5, 32, 32, 384, 384, 3), (4, 16, 16, 576, 576, 3), (4, 32, 32, 576, 384, 3), (6, 16, 16, 576, 1728, 1),
                                 (3, 16, 16, 1344, 768, 3), (2, 8, 8, 768, 576, 1)]:
    x = torch.randn(B, H, W, Cin, device=dev).to(BF)
    g = torch.randn(B, H, W, Cout, device=dev).to(BF)
    ref = ref_wgrad(g, x, ks)
    for acc in (False, True):
        dn = torch.full((Cout, ks * ks, Cin), 3.0, device=dev); do = dn.clone()
        ops.conv2d_wgrad(g, x, dn, ks, alpha=0.5, accumulate=acc)
        ops.conv2d_wgrad(g, x, do, ks, alpha=0.5, accumulate=acc, splits=-1)
        torch.cuda.synchronize()
        want = 0.5 * ref + (3.0 if acc else 0.0)
        r1, r2 = rel(dn, want), rel(do, want)
        ok = r1 < 2e-3 and not torch.isnan(dn).any().item()
        ok_all &= ok
        P(f"wgrad B{B} {H}x{W} {Cin}->{Cout} k{ks} acc={int(acc)}: pair rel={r1:.2e} single rel={r2:.2e} {'OK' if ok else 'FAIL'}")
        if not ok:
            d = (dn - want).abs()
            P("   err by tap:", [round(v, 3) for v in d.mean(dim=(0, 2)).tolist()])
            P("   err by co/16:", [round(v, 3) for v in d.mean(dim=(1, 2)).reshape(-1, 16).mean(1).tolist()])
            P("   err by ci/16:", [round(v, 3) for v in d.mean(dim=(0, 1)).reshape(-1, 16).mean(1).tolist()])
    # single split, no atomics (plain TMA store path)
    dn = torch.full((Cout, ks * ks, Cin), 3.0, device=dev)
    ops.conv2d_wgrad(g, x, dn, ks, alpha=0.5, splits=1)
    torch.cuda.synchronize()
    r = rel(dn, 0.5 * ref); ok = r < 2e-3; ok_all &= ok
    P(f"   splits=1 store path rel={r:.2e} {'OK' if ok else 'FAIL'}")
P("ALL OK" if ok_all else "SOME FAILED")
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (B, H, W, Cin, Cout, ks) in [(64, 64, 64, 192, 192, 3), (64, 32, 32, 384, 384, 3), (64, 16, 16, 576, 576, 3), (64, 8, 8, 768, 768, 3),
                                 (64, 32, 32, 576, 384, 3), (64, 64, 64, 384, 192, 3), (64, 16, 16, 960, 576, 3), (64, 16, 16, 576, 1728, 1),
                                 (128, 28, 28, 128, 128, 3), (128, 14, 14, 256, 256, 3), (128, 7, 7, 512, 512, 3), (128, 14, 14, 384, 256, 3)]:
    x = torch.randn(B, H, W, Cin, device=dev).to(BF)
    g = torch.randn(B, H, W, Cout, device=dev).to(BF)
    dw = torch.zeros(Cout, ks * ks, Cin, device=dev)
    fl = 2.0 * B * H * W * Cout * Cin * ks * ks
    tn = bench(lambda: ops.conv2d_wgrad(g, x, dw, ks))
    to = bench(lambda: ops.conv2d_wgrad(g, x, dw, ks, splits=-1))
    P(f"time wgrad B{B} {H}x{W} {Cin}->{Cout} k{ks}: pair {tn*1e3:7.1f} us {fl/tn/1e9:7.1f} TF/s | single {to*1e3:7.1f} us {fl/to/1e9:7.1f} TF/s")
