"""Timing of the fp32 GEMM at the shapes of the modulation island (CIFAR: N = 21 blocks x 256 columns, E = 256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyedm_b200 import ops
dev = torch.device("cuda:0"); ops.ensure_device(dev)
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for B in (128, 256):
    N, E = 5376, 256
    emb = torch.randn(B, E, device=dev); W = torch.randn(N, E, device=dev); lin = torch.empty(B, N, device=dev)
    d_lin = torch.randn(B, N, device=dev); g_emb = torch.empty(B, E, device=dev); gh = torch.empty(N, E, device=dev)
    for name, fn, fl in (("lin = emb W^T", lambda: ops.sgemm(emb, W, lin, B, N, E, E, E, N, False, True), 2.0 * B * N * E),
                         ("g_emb = d_lin W", lambda: ops.sgemm(d_lin, W, g_emb, B, E, N, N, E, E, False, False), 2.0 * B * N * E),
                         ("gW = d_lin^T emb", lambda: ops.sgemm(d_lin, emb, gh, N, E, B, N, E, E, True, False), 2.0 * B * N * E)):
        t = bench(fn)
        print(f"B{B} {name:18s}: {t*1e3:7.1f} us  {fl/t/1e9:6.2f} TF/s")
