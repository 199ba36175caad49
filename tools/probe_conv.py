"""GPU probe: implicit-GEMM conv forward / wgrad vs torch (run under gpurun). Writes gpurun_out/probe_conv.log."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tinyedm_b200 import _lib

os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/probe_conv.log", "w")
def P(*a):
    s = " ".join(str(x) for x in a)
    print(s); log.write(s + "\n"); log.flush()

dev = torch.device("cuda:0")
_lib.init_device(0)
st = lambda: torch.cuda.current_stream().cuda_stream

def prep_w(w):  # OIHW fp32 -> [O][kh][kw][I] bf16
    return w.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)

def conv_fwd(x, wq, ks, Cout, epi=0, alpha=1.0, raw=None, res=None, t=0.0, mod=None, drop=0.0, seed=0, bn=0):
    B, H, W, Cin = x.shape
    out = torch.empty(B, H, W, Cout, device=dev, dtype=torch.bfloat16)
    _lib.call("tedm_conv2d_forward", x.data_ptr(), wq.data_ptr(), out.data_ptr(), B, H, W, Cin, Cout, ks, epi, alpha,
              raw.data_ptr() if raw is not None else None, res.data_ptr() if res is not None else None, t,
              mod.data_ptr() if mod is not None else None, mod.shape[1] if mod is not None else 0, drop, seed, bn, st())
    return out

def rel(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()

def ref_conv(x, w):  # x NHWC bf16, w OIHW fp32 -> NHWC fp32 (bf16-rounded operands, fp32 math)
    xr = x.float().permute(0, 3, 1, 2)
    wr = w.to(torch.bfloat16).float()
    return F.conv2d(xr, wr, padding="same").permute(0, 2, 3, 1).contiguous()

torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
ok_all = True
cases = [  # B,H,W,Cin,Cout,ks,bn
    (2, 32, 32, 64, 64, 1, 0), (2, 32, 32, 64, 64, 3, 0), (4, 32, 32, 256, 256, 3, 0), (4, 16, 16, 256, 256, 3, 0),
    (5, 8, 8, 256, 256, 3, 0), (4, 16, 16, 256, 768, 1, 0), (3, 8, 8, 512, 256, 3, 0), (3, 8, 8, 512, 256, 1, 0),
    (2, 28, 28, 128, 128, 3, 0), (3, 14, 14, 256, 256, 3, 0), (5, 7, 7, 512, 512, 3, 0), (2, 64, 64, 192, 192, 3, 0),
    (4, 32, 32, 256, 256, 3, 128), (4, 32, 32, 256, 256, 3, 64), (2, 16, 16, 576, 1728, 1, 0), (2, 32, 32, 64, 16, 1, 0),
]
for (B, H, W, Cin, Cout, ks, bn) in cases:
    try:
        x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
        w = torch.randn(Cout, Cin, ks, ks, device=dev) / (Cin * ks * ks) ** 0.5
        y = conv_fwd(x, prep_w(w), ks, Cout, bn=bn)
        torch.cuda.synchronize()
        r = rel(y, ref_conv(x, w))
        ok = r < 6e-3
        ok_all &= ok
        P(f"fwd B{B} {H}x{W} {Cin}->{Cout} k{ks} bn{bn}: rel={r:.3e} {'OK' if ok else 'FAIL'}")
        if not ok:
            yr = ref_conv(x, w)
            d = (y.float() - yr).abs()
            P("   max abs diff", d.max().item(), "at", [int(v) for v in torch.unravel_index(d.argmax(), d.shape)],
              "nan:", torch.isnan(y.float()).sum().item())
            # per-row / per-col error profile
            P("   err by pixel-in-image (first 40):", [round(v, 3) for v in d.mean(dim=(0, 3)).flatten()[:40].tolist()])
            P("   err by channel (first 16):", [round(v, 3) for v in d.mean(dim=(0, 1, 2))[:16].tolist()])
    except Exception as e:
        ok_all = False
        P(f"fwd B{B} {H}x{W} {Cin}->{Cout} k{ks}: EXC {e}")

# epilogues
B, H, W, C = 4, 16, 16, 256
x = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
w = torch.randn(C, C, 3, 3, device=dev) / (C * 9) ** 0.5
mod = (torch.randn(B, C, device=dev) * 0.3 + 1).contiguous()
raw = torch.empty(B, H, W, C, device=dev, dtype=torch.bfloat16)
y = conv_fwd(x, prep_w(w), 3, C, epi=1, raw=raw, mod=mod)
yr = ref_conv(x, w)
r1 = rel(raw, yr)
z = raw.float() * mod[:, None, None, :]
r2 = rel(y, F.silu(z) / 0.596)
P(f"epi modsilu: raw rel={r1:.3e} out rel={r2:.3e}")
ok_all &= (r1 < 6e-3 and r2 < 6e-3)
yd = conv_fwd(x, prep_w(w), 3, C, epi=1, raw=raw, mod=mod, drop=0.13, seed=1234)
keep = (yd != 0).float().mean().item()
nz = yd != 0
r3 = rel(yd[nz], (y.float()[nz] / 0.87))
P(f"epi dropout: keep frac={keep:.4f} (expect 0.87) scaled rel={r3:.3e}")
ok_all &= abs(keep - 0.87) < 0.01 and r3 < 6e-3
res = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
y = conv_fwd(x, prep_w(w), 3, C, epi=2, res=res, t=0.3)
ref = (0.7 * res.float() + 0.3 * yr) / (0.7 ** 2 + 0.3 ** 2) ** 0.5
r4 = rel(y, ref)
P(f"epi mpadd: rel={r4:.3e}")
ok_all &= r4 < 6e-3

# wgrad
def wgrad(g, x, ks, alpha=1.0, splits=0):
    B, H, W, Cout = g.shape
    Cin = x.shape[3]
    dw = torch.full((Cout, ks * ks, Cin), 7.0, device=dev, dtype=torch.float32)
    _lib.call("tedm_conv2d_wgrad", g.data_ptr(), x.data_ptr(), dw.data_ptr(), B, H, W, Cin, Cout, ks, alpha, 0, splits, st())
    return dw

def ref_wgrad(g, x, ks):
    xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(False)
    Cout = g.shape[3]; Cin = x.shape[3]
    w = torch.zeros(Cout, Cin, ks, ks, device=dev, requires_grad=True)
    y = F.conv2d(xr, w, padding="same")
    (gw,) = torch.autograd.grad(y, w, g.float().permute(0, 3, 1, 2))
    return gw.permute(0, 2, 3, 1).reshape(Cout, ks * ks, Cin)

wcases = [(2, 32, 32, 64, 64, 1, 0), (2, 32, 32, 64, 128, 3, 0), (4, 32, 32, 256, 256, 3, 0), (4, 16, 16, 256, 256, 3, 0),
          (6, 8, 8, 256, 256, 3, 0), (3, 8, 8, 512, 256, 3, 0), (4, 16, 16, 256, 768, 1, 0), (2, 28, 28, 128, 128, 3, 0),
          (3, 14, 14, 256, 256, 3, 0), (5, 7, 7, 512, 512, 3, 0), (2, 64, 64, 192, 192, 3, 0), (4, 32, 32, 256, 256, 3, 1),
          (2, 32, 32, 64, 256, 1, 0)]
for (B, H, W, Cin, Cout, ks, sp) in wcases:
    try:
        x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
        g = torch.randn(B, H, W, Cout, device=dev).to(torch.bfloat16)
        dw = wgrad(g, x, ks, splits=sp)
        torch.cuda.synchronize()
        ref = ref_wgrad(g, x, ks)
        r = rel(dw, ref)
        ok = r < 2e-3
        ok_all &= ok
        P(f"wgrad B{B} {H}x{W} {Cin}->{Cout} k{ks} splits{sp}: rel={r:.3e} {'OK' if ok else 'FAIL'}")
        if not ok:
            d = (dw - ref).abs()
            P("   max abs", d.max().item(), "ref max", ref.abs().max().item(), "nan", torch.isnan(dw).sum().item())
            P("   err by tap:", [round(v, 3) for v in d.mean(dim=(0, 2)).tolist()])
            P("   err by co (first 16 of each 64):", [round(v, 3) for v in d.mean(dim=(1, 2))[::8][:32].tolist()])
            P("   err by ci (every 8th):", [round(v, 3) for v in d.mean(dim=(0, 1))[::8][:32].tolist()])
    except Exception as e:
        ok_all = False
        P(f"wgrad B{B} {H}x{W} {Cin}->{Cout} k{ks}: EXC {e}")

# timing of the dominant shapes
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for (B, H, W, Cin, Cout, ks) in [(256, 32, 32, 256, 256, 3), (256, 16, 16, 256, 256, 3), (256, 8, 8, 256, 256, 3),
                                  (256, 32, 32, 512, 256, 3), (256, 16, 16, 256, 768, 1)]:
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    w = prep_w(torch.randn(Cout, Cin, ks, ks, device=dev) / (Cin * ks * ks) ** 0.5)
    g = torch.randn(B, H, W, Cout, device=dev).to(torch.bfloat16)
    fl = 2.0 * B * H * W * Cin * Cout * ks * ks
    for bn in ([0, 128] if Cout == 256 else [0]):
        ms = bench(lambda: conv_fwd(x, w, ks, Cout, bn=bn))
        P(f"time fwd B{B} {H}x{W} {Cin}->{Cout} k{ks} bn{bn}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
    ms = bench(lambda: wgrad(g, x, ks))
    P(f"time wgrad B{B} {H}x{W} {Cin}->{Cout} k{ks}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
    # cuDNN reference (channels_last bf16)
    xc = x.permute(0, 3, 1, 2); wc = torch.randn(Cout, Cin, ks, ks, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ms = bench(lambda: F.conv2d(xc, wc, padding="same"))
    P(f"time cudnn-NHWC fwd same shape: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")

P("ALL OK" if ok_all else "SOME FAILED")
sys.exit(0 if ok_all else 1)
