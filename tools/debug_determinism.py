import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tinyedm_b200 as T
from tests.helpers import SMALL, build_modules, rel, small_params
dev = torch.device("cuda:0")
dp, ep, _ = small_params()
den, emb_m, _ = build_modules(SMALL, dp, ep, None, dev)
with torch.no_grad(): den.gain_out.fill_(1.0)
torch.manual_seed(0)
B = 8
noisy = torch.randn(B, 3, 16, 16, device=dev)
sigma = torch.rand(B, device=dev) + 0.3
emb_m.eval()
_, e = emb_m(sigma, torch.randint(0, 10, (B,), device=dev))
e = e.detach()
eng = den.engine
for training in (False, True):
    outs = []
    junk = []
    for it in range(4):
        taps = {}
        D, ctx = eng.forward(noisy, sigma, e, training=training, save=True, taps=taps)
        torch.cuda.synchronize()
        outs.append(({k: v.float().clone() for k, v in taps.items()}, D.clone(), ctx))
        junk.append(torch.full((1 << 20,), float("nan"), device=dev))   # perturb the allocator + poison freed memory
        if it % 2: junk = []
    print("training" if training else "eval")
    for it in range(1, 4):
        diffs = [(k, rel(outs[it][0][k], outs[0][0][k])) for k in outs[0][0]]
        first = next(((k, d) for k, d in diffs if d > 0), None)
        print(f"  run {it} vs run 0: D rel {rel(outs[it][1], outs[0][1]):.3e}; first differing tap: {first}")
    # saved tensors of block 0 / 1
    for bi in range(3):
        for k in ("x", "a", "raw", "h"):
            a, b = outs[1][2]["blocks"][bi].get(k), outs[0][2]["blocks"][bi].get(k)
            if a is not None:
                print(f"    block {bi} saved {k}: rel {rel(a.float(), b.float()):.3e} nan={int(torch.isnan(a.float()).sum())}")
