"""GPU probe: CTA-pair conv kernel (conv_pair.cu) vs the single-CTA kernel (conv_gemm.cu) and vs torch, every epilogue,
plus timings of the CIFAR shapes. Run under gpurun; writes gpurun_out/probe_pair.log."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tinyedm_b200 import _lib, ops
from tinyedm_b200.ops import EPI_PLAIN, EPI_MODSILU, EPI_AXPBY, EPI_MODSILU_BWD, EPI_SILU_BWD

os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/probe_pair.log", "w")
def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True); log.write(s + "\n"); log.flush()

dev = torch.device("cuda:0")
ops.ensure_device(dev)
torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
BF = torch.bfloat16

def rel(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()

def prep_w(w):
    return w.permute(0, 2, 3, 1).contiguous().to(BF).reshape(w.shape[0], -1)

def ref_conv(x, w):
    return F.conv2d(x.float().permute(0, 3, 1, 2), w.to(BF).float(), padding="same").permute(0, 2, 3, 1).contiguous()

ok_all = True
def check(name, new, old, tol=1e-5, ref=None, reftol=6e-3):
    global ok_all
    torch.cuda.synchronize()
    r = rel(new, old)
    msg = f"{name}: new-vs-old rel={r:.2e}"
    ok = r <= tol and not torch.isnan(new.float()).any().item()
    if ref is not None:
        rr = rel(new, ref)
        msg += f" new-vs-torch rel={rr:.2e}"
        ok = ok and rr < reftol
    ok_all &= ok
    P(msg, "OK" if ok else "FAIL")
    if not ok:
        d = (new.float() - old.float()).abs()
        if d.dim() == 4:
            P("   err by image:", [round(v, 4) for v in d.mean(dim=(1, 2, 3)).tolist()][:16])
            P("   err by row:", [round(v, 4) for v in d.mean(dim=(0, 2, 3)).tolist()][:32])
            P("   err by col:", [round(v, 4) for v in d.mean(dim=(0, 1, 3)).tolist()][:32])
            P("   err by channel/16:", [round(v, 4) for v in d.mean(dim=(0, 1, 2)).reshape(-1, 16).mean(1).tolist()])

shapes = [  # B,H,W,Cin,Cout,ks
    (4, 32, 32, 256, 256, 3), (3, 32, 32, 256, 256, 3), (5, 16, 16, 256, 256, 3), (7, 8, 8, 256, 256, 3),
    (4, 16, 16, 256, 768, 1), (3, 8, 8, 512, 256, 3), (2, 16, 16, 256, 512, 3), (2, 28, 28, 128, 128, 3),
    (3, 14, 14, 256, 256, 3), (5, 7, 7, 512, 512, 3), (2, 64, 64, 192, 192, 3), (2, 32, 32, 384, 384, 3),
    (1, 32, 32, 64, 256, 1), (300, 8, 8, 256, 256, 3),
]
for (B, H, W, Cin, Cout, ks) in shapes:
    tag = f"B{B} {H}x{W} {Cin}->{Cout} k{ks}"
    try:
        x = torch.randn(B, H, W, Cin, device=dev).to(BF)
        w = torch.randn(Cout, Cin, ks, ks, device=dev) / (Cin * ks * ks) ** 0.5
        wq = prep_w(w)
        yr = ref_conv(x, w)
        bn_old = 256 if Cout >= 256 else 128
        # plain
        check(f"plain    {tag}", ops.conv2d(x, wq, ks, Cout, alpha=0.7), ops.conv2d(x, wq, ks, Cout, alpha=0.7, block_n=bn_old), ref=0.7 * yr)
        # mp_add
        res = torch.randn(B, H, W, Cout, device=dev).to(BF)
        kw = dict(epi=EPI_AXPBY, alpha=0.4, beta=0.9, res=res)
        check(f"axpby    {tag}", ops.conv2d(x, wq, ks, Cout, **kw), ops.conv2d(x, wq, ks, Cout, block_n=bn_old, **kw), ref=0.4 * yr + 0.9 * res.float())
        # modulation * silu * dropout (+ raw copy)
        mod = (torch.randn(B, Cout + 64, device=dev) * 0.3 + 1).contiguous()
        raw_n = torch.zeros(B, H, W, Cout, device=dev, dtype=BF); raw_o = torch.zeros_like(raw_n)
        kw = dict(epi=EPI_MODSILU, mod=mod, mod_off=64, drop_p=0.13, seed=77)
        yn = ops.conv2d(x, wq, ks, Cout, raw=raw_n, **kw)
        yo = ops.conv2d(x, wq, ks, Cout, raw=raw_o, block_n=bn_old, **kw)
        check(f"modsilu  {tag}", yn, yo)
        check(f"  raw    {tag}", raw_n, raw_o, ref=yr)
        # adjoint of modulation*silu*dropout
        rawt = torch.randn(B, H, W, Cout, device=dev).to(BF)
        dm_n = torch.zeros(B, Cout + 64, device=dev); dm_o = torch.zeros_like(dm_n)
        kw = dict(epi=EPI_MODSILU_BWD, alpha=0.5, aux=rawt, mod=mod, mod_off=64, drop_p=0.13, seed=77)
        yn = ops.conv2d(x, wq, ks, Cout, d_mod=dm_n, **kw)
        yo = ops.conv2d(x, wq, ks, Cout, d_mod=dm_o, block_n=bn_old, **kw)
        check(f"modsilu' {tag}", yn, yo)
        check(f"  d_mod  {tag}", dm_n, dm_o, tol=2e-3)
        # adjoint of mp_silu (+ residual share, + accumulate)
        xs = torch.randn(B, H, W, Cout, device=dev).to(BF)
        kw = dict(epi=EPI_SILU_BWD, alpha=1.1, aux=xs, res=res, beta=0.6)
        check(f"silu'    {tag}", ops.conv2d(x, wq, ks, Cout, **kw), ops.conv2d(x, wq, ks, Cout, block_n=bn_old, **kw))
        kw = dict(epi=EPI_SILU_BWD, alpha=1.1, aux=xs)
        o_n = torch.randn(B, H, W, Cout, device=dev).to(BF); o_o = o_n.clone()
        ops.conv2d(x, wq, ks, Cout, out=o_n, accumulate_out=True, **kw)
        ops.conv2d(x, wq, ks, Cout, out=o_o, accumulate_out=True, block_n=bn_old, **kw)
        check(f"silu'+=  {tag}", o_n, o_o)
        if Cout <= 256:
            nrm = (torch.rand(B, H, W, device=dev) + 0.5).contiguous()
            kw = dict(epi=EPI_SILU_BWD, alpha=1.1, aux=xs, res=res, beta=0.6, nrm=nrm)
            check(f"silu'pn  {tag}", ops.conv2d(x, wq, ks, Cout, **kw), ops.conv2d(x, wq, ks, Cout, block_n=256, **kw), tol=3e-3)
            o_n = torch.randn(B, H, W, Cout, device=dev).to(BF); o_o = o_n.clone()
            ops.conv2d(x, wq, ks, Cout, out=o_n, accumulate_out=True, **kw)
            ops.conv2d(x, wq, ks, Cout, out=o_o, accumulate_out=True, block_n=256, **kw)
            check(f"silu'pn+={tag}", o_n, o_o, tol=3e-3)
    except Exception as e:
        ok_all = False
        P(f"{tag}: EXC {e}")
        break
P("ALL OK" if ok_all else "SOME FAILED")

# ---- timings: CIFAR training shapes, B = 256 ----
def bench(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

if ok_all or os.environ.get("PROBE_TIME_ANYWAY"):
    for (B, H, W, Cin, Cout, ks) in [(256, 32, 32, 256, 256, 3), (256, 16, 16, 256, 256, 3), (256, 8, 8, 256, 256, 3),
                                     (256, 32, 32, 512, 256, 3), (256, 16, 16, 256, 768, 1), (256, 16, 16, 256, 256, 1),
                                     (128, 32, 32, 256, 256, 3), (128, 16, 16, 256, 256, 3), (128, 8, 8, 256, 256, 3)]:
        x = torch.randn(B, H, W, Cin, device=dev).to(BF)
        w = torch.randn(Cout, Cin, ks, ks, device=dev) / (Cin * ks * ks) ** 0.5
        wq = prep_w(w)
        res = torch.randn(B, H, W, Cout, device=dev).to(BF)
        xs = torch.randn(B, H, W, Cout, device=dev).to(BF)
        mod = (torch.randn(B, Cout, device=dev) * 0.3 + 1).contiguous()
        raw = torch.empty(B, H, W, Cout, device=dev, dtype=BF)
        out = torch.empty(B, H, W, Cout, device=dev, dtype=BF)
        dm = torch.zeros(B, Cout, device=dev)
        nrm = (torch.rand(B, H, W, device=dev) + 0.5).contiguous()
        fl = 2.0 * B * H * W * Cout * Cin * ks * ks
        variants = {
            "plain": dict(),
            "axpby": dict(epi=EPI_AXPBY, alpha=0.4, beta=0.9, res=res),
            "modsilu": dict(epi=EPI_MODSILU, mod=mod, drop_p=0.13, seed=5, raw=raw),
            "modsilu'": dict(epi=EPI_MODSILU_BWD, aux=xs, mod=mod, drop_p=0.13, seed=5, d_mod=dm),
            "silu'": dict(epi=EPI_SILU_BWD, aux=xs, res=res, beta=0.6),
        }
        if Cout <= 256:
            variants["silu'pn"] = dict(epi=EPI_SILU_BWD, aux=xs, res=res, beta=0.6, nrm=nrm)
        for name, kw in variants.items():
            t_new = bench(lambda: ops.conv2d(x, wq, ks, Cout, out=out, **kw))
            t_old = bench(lambda: ops.conv2d(x, wq, ks, Cout, out=out, block_n=256 if "pn" in name else -1, **kw)) if False else None
            t_old = bench(lambda: ops.conv2d(x, wq, ks, Cout, out=out, block_n=256, **kw))
            P(f"time B{B} {H}x{W} {Cin}->{Cout} k{ks} {name:9s}: pair {t_new*1e3:7.1f} us {fl/t_new/1e9:7.1f} TF/s | single(bn256) {t_old*1e3:7.1f} us {fl/t_old/1e9:7.1f} TF/s")
