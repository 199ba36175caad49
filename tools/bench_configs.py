"""Secondary configs of BASELINE.json (parity-test cases, not bench lines): eager training-step and network-evaluation
times for the MNIST and ImageNet-512-latent architectures, so that nothing on their paths is pathologically slow."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tinyedm_b200 as T
from tinyedm_b200.configs import IMAGENET, MNIST, build_edm
dev = torch.device("cuda:0")
out = {}
for name, cfg, B, gflop in (("mnist", MNIST, 128, 20.11), ("imagenet_latent", IMAGENET, int(os.environ.get("IN_B", "64")), 192.9)):
    torch.manual_seed(0)
    model = build_edm(cfg).to(dev).train()
    with torch.no_grad(): model.denoiser.gain_out.fill_(1.0)
    opt = model.configure_optimizers()["optimizer"]
    for g in opt.param_groups: g["lr"] = 1e-5
    C, H, W = cfg["image"]
    x = (0.5 * torch.randn(B, C, H, W, device=dev)).clamp(-1, 1)
    y = torch.randint(0, cfg["embedding"]["num_classes"], (B,), device=dev)
    def step():
        opt.zero_grad(set_to_none=True); loss = model.training_step((x, y), 0); loss.backward(); opt.step(); return loss
    for _ in range(3): l = step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 5
    for _ in range(n): l = step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    model.eval()
    sig = torch.full((B,), 1.5, device=dev)
    with torch.no_grad():
        for _ in range(3): model(x, sig, y)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): model(x, sig, y)
        torch.cuda.synchronize(); dn = (time.perf_counter() - t0) / n
    # 32-step Heun sampling through the package's solver (whole-trajectory CUDA graph), device noise in -> device images out
    solver = T.DeterministicSolver(num_steps=32)
    x0 = torch.randn(B, C, H, W, device=dev)
    lab = y.view(B, 1)
    with torch.no_grad():
        solver.solve(model, x0, lab)                      # eager warm-up + capture + first replay
        torch.cuda.synchronize(); t0 = time.perf_counter()
        img = solver.solve(model, x0, lab)
        torch.cuda.synchronize(); ds = time.perf_counter() - t0
    assert bool(torch.isfinite(img).all())
    out[name] = {"batch": B, "sample_ms_per_solve": ds * 1e3, "sample_img_s": B / ds, "train_ms": dt * 1e3, "train_img_s": B / dt, "train_tflops": 3 * gflop * B / dt / 1e3,
                 "nfe_ms": dn * 1e3, "nfe_tflops": gflop * B / dn / 1e3, "loss": float(l.detach()),
                 "params_M": sum(p.numel() for p in model.parameters()) / 1e6}
    print(name, json.dumps(out[name]), flush=True)
    del model, opt
    torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/bench_configs.json", "w"))
