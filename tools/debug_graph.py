import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tinyedm_b200 as T
from tests.helpers import SMALL, build_modules, rel, small_params
dev = torch.device("cuda:0")
dp, ep, _ = small_params()
den, emb_m, _ = build_modules(SMALL, dp, ep, None, dev)
m = T.EDM(diffuser=T.Diffuser(-1.2, 1.2), embedding=emb_m, denoiser=den, use_ema=False, use_uncertainty=False,
          steady_steps=1, rampup_steps=1, scheduler_interval="step").train()
print("dropout", den.dropout_rate, "cfg", SMALL["denoiser"])
with torch.no_grad(): m.denoiser.gain_out.fill_(1.0)
torch.manual_seed(0)
B = 8
clean = (0.5 * torch.randn(B, 3, 16, 16, device=dev)).clamp(-1, 1)
labels = torch.randint(0, 10, (B,), device=dev)
fs = torch.exp(torch.randn(B, device=dev) * 1.2 - 1.2); fn = torch.randn_like(clean)
m.diffuser.forward = lambda x: (x + fs.view(-1, 1, 1, 1) * fn, fs)
params = dict(m.named_parameters())
def fb():
    for p in params.values(): p.grad = None
    loss = m.training_step((clean, labels), 0)
    loss.backward()
    return loss
eng = m.denoiser.engine
def snap():
    torch.cuda.synchronize()
    return {"loss": None, "grads": {n: p.grad.clone() for n, p in params.items() if p.grad is not None},
            "ghat": eng.bank._ghat_flat.clone(), "sg": eng.bank._ghat_flat[:len(eng.blocks) + 1].clone()}
for _ in range(2): fb()
e1 = snap(); l1 = float(fb().detach()); e2 = snap()
print("eager vs eager ghat", rel(e2["ghat"], e1["ghat"]), "sg", rel(e2["sg"], e1["sg"]))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    lg = fb()
for it in range(3):
    g.replay(); r = snap()
    print(f"replay {it}: loss eager {l1:.6f} graph {float(lg):.6f}; ghat rel {rel(r['ghat'], e1['ghat']):.3e} sg rel {rel(r['sg'], e1['sg']):.3e}")
    errs = {k: rel(r["grads"][k], e1["grads"][k]) for k in e1["grads"] if float(e1["grads"][k].norm()) > 0}
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print("   top grad diffs:", [(k, f"{v:.2e}") for k, v in top])
    # per-slot ghat diffs
    worst = []
    for s in eng.bank.slots:
        if s.ghat is not None:
            o = s.ghat.storage_offset(); n = s.ghat.numel()
            a, b = r["ghat"][o:o + n], e1["ghat"][o:o + n]
            if float(b.norm()) > 0: worst.append((rel(a, b), s.name))
    worst.sort(reverse=True)
    print("   top ghat diffs:", [(n, f"{v:.2e}") for v, n in worst[:6]])
fb(); e3 = snap()
print("eager after graph vs eager before: ghat", rel(e3["ghat"], e1["ghat"]))
