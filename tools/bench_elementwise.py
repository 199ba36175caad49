"""GB/s of the HBM-bound kernels at the CIFAR training shapes (B=256) against their ALGORITHMIC bytes
(every distinct input read once + every output written once, storage dtype).
Run under gpurun; writes gpurun_out/bench_elementwise.log. Peak: MEASURED_PEAKS.json hbm_gbs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyedm_b200 import ops
from tinyedm_b200.ops import RESAMPLE_DOWN, RESAMPLE_NONE, RESAMPLE_UP

os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/bench_elementwise.log", "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); log.write(s + "\n"); log.flush()
dev = torch.device("cuda:0"); ops.ensure_device(dev); BF = torch.bfloat16; F32 = torch.float32
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6551.7
flush_buf = torch.empty(256 << 20, device=dev, dtype=torch.uint8)   # > L2 (126 MB)

ITERS = int(os.environ.get("EW_ITERS", 10))   # EW_ITERS=1 under ncu: one cold launch per case
def bench(fn, n=ITERS):
    for _ in range(2 if n > 1 else 0):
        fn()
    ts = []
    for _ in range(n):
        flush_buf.zero_()                       # evict the operands from L2 between iterations
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

def report(name, nbytes, fn):
    t = bench(fn)
    gbs = nbytes / t / 1e6
    P(f"{name:58s} {nbytes/1e6:8.1f} MB {t*1e3:8.1f} us {gbs:8.1f} GB/s  {100*gbs/PEAK:5.1f}% of {PEAK:.0f}")

B = int(os.environ.get("EW_B", 256))
def rb(*shape): return torch.randn(*shape, device=dev).to(BF)
for (H, C) in [(32, 256), (16, 256), (8, 256)]:
    W = H
    n = B * H * W * C * 2       # bytes of one (B,H,W,C) bf16 tensor
    x = rb(B, H, W, C); s = rb(B, H, W, C); gain = torch.rand(B, C, device=dev)
    a = torch.empty(100 << 20, device=dev); b_ = torch.empty_like(a)
    report(f"[{H}x{H}] torch copy_ 100M fp32 (reference)", 800e6 * 1.048576, lambda: b_.copy_(a))
    del a, b_
    report(f"[{H}x{H}] block_prep fwd pixelnorm+silu (enc)  r C, w 2C", 3 * n, lambda: ops.block_prep(x, pixelnorm=True, want_nrm=True))
    report(f"[{H}x{H}] block_prep fwd silu only (dec)       r C, w C", 2 * n, lambda: ops.block_prep(x, want_x=False))
    report(f"[{H}x{H}] block_prep fwd concat*gain+silu      r 2C, w 4C", 6 * n, lambda: ops.block_prep(x, skip=s, gain=gain))
    if H > 8:
        report(f"[{H}x{H}] block_prep fwd down+pixelnorm+silu   r C, w C/2", 1.5 * n, lambda: ops.block_prep(x, resample=RESAMPLE_DOWN, pixelnorm=True, want_nrm=True))
    if H < 32:
        report(f"[{H}x{H}] block_prep fwd up+silu               r C, w 8C", 9 * n, lambda: ops.block_prep(x, resample=RESAMPLE_UP))
    # backward: split of the concatenated gradient (+ ScaleLong mean-gradient, accumulate into a pending skip gradient)
    gcat = rb(B, H, W, 2 * C); g_in = rb(B, H, W, C); g_skip = torch.empty(B, H, W, C, device=dev, dtype=BF)
    d_mean = torch.randn(B, C, device=dev)
    kw = dict(g_res=gcat, beta=1.0, g_a=None, x=None, nrm=None, gain=gain, d_mean=d_mean, g_in=g_in, g_skip=g_skip,
              accumulate_skip=False, B=B, Hin=H, Win=W, C1=C, C2=C, resample=RESAMPLE_NONE, pixelnorm=False)
    report(f"[{H}x{H}] block_prep bwd split                 r 2C, w 2C", 4 * n, lambda: ops.block_prep_backward(accumulate_in=False, **kw))
    report(f"[{H}x{H}] block_prep bwd split (+= into g_in)  r 3C, w 2C", 5 * n, lambda: ops.block_prep_backward(accumulate_in=True, **kw))
    mean = torch.zeros(B, C, device=dev)
    report(f"[{H}x{H}] channel_dot mean (ScaleLong input)   r C", n, lambda: ops.channel_dot(s, None, mean, C, 0, 1.0 / (H * W)))
    report(f"[{H}x{H}] channel_dot gain gradient            r 2C", 2 * n, lambda: ops.channel_dot(gcat, s, mean, C, C, 1.0))
    if H < 32:
        gx = rb(B, 2 * H, 2 * W, C); gi = torch.empty(B, H, W, C, device=dev, dtype=BF)
        kw2 = dict(g_res=gx, beta=1.0, g_a=None, x=None, nrm=None, gain=None, d_mean=None, g_in=gi, g_skip=None,
                   accumulate_in=False, accumulate_skip=False, B=B, Hin=H, Win=W, C1=C, C2=0, resample=RESAMPLE_UP, pixelnorm=False)
        report(f"[{H}x{H}] block_prep bwd adjoint of upsample   r 4C, w C", 5 * n, lambda: ops.block_prep_backward(**kw2))
    del x, s, gcat, g_in, g_skip

# image-sized and parameter-sized kernels
from tinyedm_b200 import configs
from tinyedm_b200.networks import Denoiser
den = Denoiser(**configs.CIFAR10["denoiser"]).to(dev)
if den is not None:
    eng = den.engine
    eng._ensure_device(dev)
    bank = eng.bank
    Pn = sum(s.param.numel() for s in bank.slots)
    bank.ensure_grad_buffers()
    report("weight_prep fwd (train: rewrite + 2 bf16 layouts)  r 4P, w 4P+2P+2P", 12 * Pn, lambda: (bank.invalidate(), bank.prepare(True)))
    report("weight_prep fwd (eval)                           r 4P, w 2P+2P", 8 * Pn, lambda: (bank.invalidate(), bank._build_table() if bank._table is None else None, ops.weight_prep_forward(bank._table, len(bank.slots), bank.total_groups, False)))
    report("weight_prep bwd                                  r 4P+4P, w 4P", 12 * Pn, lambda: bank.backward())

# conv_out (+ output preconditioning) and its adjoint at the CIFAR training shape
Hh = 32; C = 256; Co = 3
x = rb(B, Hh, Hh, C); w = (torch.randn(Co, C, device=dev) / 16).to(BF); gain_out = torch.tensor(0.8, device=dev)
noisy = torch.randn(B, Co, Hh, Hh, device=dev); sigma = torch.rand(B, device=dev) + 0.1
n = B * Hh * Hh * C * 2; ni = B * Co * Hh * Hh * 4
report("[32x32] conv_out fwd + precond                        r C, r/w image", n + 3 * ni, lambda: ops.conv_out_forward(x, w, gain_out, noisy, sigma, 0.5, True))
D, f_raw = ops.conv_out_forward(x, w, gain_out, noisy, sigma, 0.5, True)
g_D = torch.randn_like(D); g_w = torch.zeros(Co, C, device=dev); g_g = torch.zeros((), device=dev)
report("[32x32] conv_out bwd                                  r C, w C, r image", 2 * n + 2 * ni, lambda: ops.conv_out_backward(g_D, f_raw, x, w, gain_out, sigma, 0.5, g_w, g_g))

# round 2: the kernels added on the HBM side — q, k, v normalisation of the generic attention path, the fused Philox
# diffuser / input block, the fused Adam + EMA step
for (name, Bq, S, heads, hd) in [("MNIST 14x14 hd 64", 128, 196, 4, 64), ("MNIST 7x7 hd 128", 128, 49, 4, 128),
                                 ("ImageNet-latent 16x16 hd 144", 176, 256, 4, 144), ("ImageNet-latent 8x8 hd 192", 176, 64, 4, 192)]:
    Hs = int(S ** 0.5)
    qkv = rb(Bq, Hs, Hs, 3 * heads * hd)
    nb = qkv.numel() * 2
    report(f"qkv_normalize {name:32s} r 3C, w 3C (+norms)", 2 * nb + Bq * S * 3 * heads * 4, lambda: ops.qkv_normalize(qkv, heads))
    del qkv
for (name, Bd, Ci, Hd) in [("CIFAR B=256 3x32x32", 256, 3, 32), ("MNIST B=128 1x28x28", 128, 1, 28), ("ImageNet-latent B=176 4x64x64", 176, 4, 64)]:
    clean = torch.randn(Bd, Ci, Hd, Hd, device=dev)
    step = torch.tensor([1234, 0], dtype=torch.int64, device=dev)
    ni = clean.numel() * 4
    report(f"diffuse_philox + input block {name:30s} r img, w img + 128 B/pixel", 2 * ni + Bd * Hd * Hd * 128,
           lambda: ops.diffuse_philox(clean, step, -1.2, 1.2, 0.5))
    del clean
import tinyedm_b200 as T
ps = [torch.nn.Parameter(torch.randn(35_600_000 // 8, device=dev)) for _ in range(8)]
for p_ in ps: p_.grad = torch.randn_like(p_)
opt = T.FusedAdamEMA(ps, lr=1e-3, ema_length=0.13)
opt.step()
report("fused Adam + EMA, 35.6 M parameters                        r p,g,m,v,e  w p,m,v,e (36 B/param)", 36 * 35_600_000, lambda: opt.step())
