"""GPU probe: weight_prep forward (training / eval) and backward over the weight banks of the three configs, against their
algorithmic bytes. TEDM_WPB selects the backward variant under test."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyedm_b200 import configs, ops
from tinyedm_b200.networks import Denoiser
dev = torch.device("cuda:0"); ops.ensure_device(dev)
PEAK = 6551.7
flush_buf = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
def bench(fn, n=7):
    fn(); fn()
    ts = []
    for _ in range(n):
        flush_buf.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]
for name, cfg in (("cifar10", configs.CIFAR10), ("mnist", configs.MNIST), ("imagenet", configs.IMAGENET)):
    den = Denoiser(**cfg["denoiser"]).to(dev)
    eng = den.engine; eng._ensure_device(dev); bank = eng.bank
    Pn = sum(s.param.numel() for s in bank.slots)
    bank.ensure_grad_buffers()
    bank.prepare(True)
    for label, nbytes, fn in (("fwd train", 12 * Pn, lambda: (bank.invalidate(), bank.prepare(True))),
                              ("fwd eval ", 8 * Pn, lambda: ops.weight_prep_forward(bank._table, len(bank.slots), bank.total_groups, False)),
                              ("bwd      ", 12 * Pn, lambda: bank.backward())):
        t = bench(fn)
        print(f"{name:9s} P={Pn/1e6:6.1f}M weight_prep {label} {nbytes/1e6:8.1f} MB {t*1e3:8.1f} us {nbytes/t/1e6:7.1f} GB/s {100*nbytes/t/1e6/PEAK:5.1f}% (TEDM_WPB={os.environ.get('TEDM_WPB','0')})", flush=True)
    del den, eng, bank
    torch.cuda.empty_cache()
