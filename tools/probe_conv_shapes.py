"""GPU probe: achieved TFLOP/s of the conv (forward / data-gradient epilogues) and weight-gradient kernels on the shapes of
the ImageNet-latent (micro-batch 176), MNIST (128) and CIFAR (256 / 128) configs, one launch flavour per line, L2-cold
(the inputs of a launch are re-used only after > 126 MB of other traffic). Writes gpurun_out/probe_conv_shapes.txt."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyedm_b200 import ops
from tinyedm_b200.ops import EPI_PLAIN, EPI_MODSILU, EPI_AXPBY, EPI_MODSILU_BWD, EPI_SILU_BWD

dev = torch.device("cuda:0")
ops.ensure_device(dev)
BF = torch.bfloat16
os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/probe_conv_shapes.txt", "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); log.write(s + "\n"); log.flush()

def timeit(fn, n=8):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3   # us

CONFIGS = {
    "imagenet B=176": [(176, 64, 192, 192, 3), (176, 64, 384, 192, 3), (176, 64, 384, 192, 1), (176, 32, 384, 384, 3), (176, 32, 576, 384, 3),
                       (176, 32, 192, 384, 1), (176, 16, 576, 576, 3), (176, 16, 1152, 576, 3), (176, 16, 576, 1728, 1), (176, 16, 576, 576, 1),
                       (176, 8, 768, 768, 3), (176, 8, 1536, 768, 3), (176, 8, 768, 2304, 1), (176, 8, 768, 768, 1)],
    "mnist B=128": [(128, 28, 128, 128, 3), (128, 28, 256, 128, 3), (128, 14, 256, 256, 3), (128, 14, 512, 256, 3), (128, 14, 256, 768, 1),
                    (128, 7, 512, 512, 3), (128, 7, 1024, 512, 3), (128, 7, 512, 1536, 1)],
    "cifar B=256": [(256, 32, 256, 256, 3), (256, 16, 256, 256, 3), (256, 8, 256, 256, 3), (256, 16, 512, 256, 3), (256, 16, 256, 768, 1)],
    "cifar sampling B=128": [(128, 32, 256, 256, 3), (128, 16, 256, 256, 3), (128, 8, 256, 256, 3), (128, 32, 512, 256, 3)],
}
only = os.environ.get("ONLY")
for name, shapes in CONFIGS.items():
    if only and only not in name: continue
    P(f"== {name}: us and TFLOP/s per launch: plain | modsilu | axpby | dgrad modsilu-adjoint | dgrad silu-adjoint | wgrad")
    for (B, H, Cin, Cout, ks) in shapes:
        torch.manual_seed(0)
        nbuf = 3                                          # rotate operands so that a launch does not find its inputs in L2
        xs = [torch.randn(B, H, H, Cin, device=dev).to(BF) for _ in range(nbuf)]
        w = (torch.randn(Cout, ks * ks * Cin, device=dev) / (ks * ks * Cin) ** 0.5).to(BF)
        wd = (torch.randn(Cin, ks * ks * Cout, device=dev) / (ks * ks * Cout) ** 0.5).to(BF)
        res = torch.randn(B, H, H, Cout, device=dev).to(BF)
        gs = [torch.randn(B, H, H, Cout, device=dev).to(BF) for _ in range(nbuf)]
        xin = torch.randn(B, H, H, Cin, device=dev).to(BF)
        mod = torch.rand(B, Cout, device=dev) + 0.5
        modi = torch.rand(B, Cin, device=dev) + 0.5
        d_mod = torch.zeros(B, Cin, device=dev)
        out = torch.empty(B, H, H, Cout, device=dev, dtype=BF)
        outi = torch.empty(B, H, H, Cin, device=dev, dtype=BF)
        dw = torch.zeros(Cout, ks * ks, Cin, device=dev)
        flop = 2.0 * B * H * H * Cin * Cout * ks * ks
        i = [0]
        def nx(lst):
            i[0] += 1
            return lst[i[0] % nbuf]
        t = {}
        t["plain"] = timeit(lambda: ops.conv2d(nx(xs), w, ks, Cout, out=out))
        t["modsilu"] = timeit(lambda: ops.conv2d(nx(xs), w, ks, Cout, epi=EPI_MODSILU, mod=mod, out=out))
        t["axpby"] = timeit(lambda: ops.conv2d(nx(xs), w, ks, Cout, epi=EPI_AXPBY, alpha=0.4, beta=0.9, res=res, out=out))
        t["d-modsilu"] = timeit(lambda: ops.conv2d(nx(gs), wd, ks, Cin, epi=EPI_MODSILU_BWD, aux=xin, mod=modi, d_mod=d_mod, out=outi))
        t["d-silu"] = timeit(lambda: ops.conv2d(nx(gs), wd, ks, Cin, epi=EPI_SILU_BWD, aux=xin, res=xin, beta=0.5, out=outi))
        t["wgrad"] = timeit(lambda: ops.conv2d_wgrad(nx(gs), nx(xs), dw, ks, accumulate=True))
        P(f"{H:3d}x{H:<3d} {Cin:5d}->{Cout:<5d} k{ks}  " + " | ".join(f"{k} {v:7.1f} {flop / v / 1e6:6.0f}" for k, v in t.items()))
        del xs, gs, res, xin, out, outi, w, wd
        torch.cuda.empty_cache()
