#!/usr/bin/env python
"""Headline benchmark of the tinyedm_b200 hot path (contract: see the task's bench.py section / DESIGN.md §measurement).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

Headline workload (BASELINE.json configs[1], experiments/conf/cifar10.yaml): CIFAR-10 35.6 M unconditional EDM2 training,
per-GPU batch 256, bf16 network / fp32 master weights, dropout 0.13, fused Adam(lr 0.02 x ramp-up factor) + EMA(0.13); a
"step" is noise -> embedding -> denoiser forward -> loss -> backward -> gradient all-reduce (N>1) -> optimiser, on
synthetic images 0.5*N(0,1) clipped to [-1,1] with random-init weights (gain_out set to 1 so the network is not hidden).
The same line carries the second half of the metric — class-conditional 32-step Heun sampling (configs[2], batch 128 per
GPU, 63 network evaluations, batch-sharded) — and, under "configs", the other BASELINE.json configurations measured the
same way: MNIST training (configs[0]), CIFAR training with the uncertainty head, ImageNet-512
latent training with micro-batch 176 x 3 accumulation (configs[3], imagenet.yaml:7,14) and ImageNet-latent 32-step
sampling (configs[4]; the 50 000-image sweep is reported as img/s on a bounded sample plus the implied minutes).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "CIFAR-10 EDM2 train img/s & 32-step Heun sample img/s at 1/2/4/8 B200"
TRAIN_BATCH = 256
SAMPLE_BATCH = 128
SAMPLE_STEPS = 32
EMA_GAMMA = 0.999                   # CPU arm's EMA decay (the value does not change the work: one lerp over every parameter)
LR_RAMP = 1e-3                      # the schedule's ramp-up region (edm.py:306-317): keeps random-init training tame over the run
# forward GFLOP per image (conv3x3 + conv1x1 + attention + linear), per-GPU batch, accumulation
WORKLOADS = {
    "cifar": dict(fwd_gflop=27.001, batch=256, sample_batch=128, accumulate=1),
    "mnist": dict(fwd_gflop=20.105, batch=128, sample_batch=128, accumulate=1),
    # (sampling batch of the 50 000-image sweep is free: 128 per GPU measured 72.7 img/s against 69.8 at 64 and 73.5 at 176,
    #  tools/probe_sampling_batch.py)
    "imagenet": dict(fwd_gflop=192.886, batch=176, sample_batch=128, accumulate=3),
}
FWD_GFLOP_PER_IMG = WORKLOADS["cifar"]["fwd_gflop"]
# dominant kernel: conv_pair_kernel, 3x3 256->256 at 32x32, per-GPU batch 256
DOM = dict(B=TRAIN_BATCH, H=32, W=32, Cin=256, Cout=256, k=3)
DOM_FLOP = 2.0 * DOM["B"] * DOM["H"] * DOM["W"] * DOM["Cin"] * DOM["Cout"] * DOM["k"] ** 2


def peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        return {}


def dominant_kernel_traffic() -> dict:
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this build's kernel
    (profiles/conv_pair_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep): measured under ncu, cited here."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "conv_pair_traffic.json")))
    except Exception:  # noqa: BLE001
        return {"traffic_bytes": None, "source": "no capture committed"}


# ---------------------------------------------------------------------------------------------------------
# clocks under load (nvidia-smi sampled DURING the timed region)
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark(self) -> int:
        """Index of the next sample: call at both ends of the timed region."""
        return len(self.lines)

    def summarise(self, first: int = 0, last: int | None = None) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[first:last]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power)}

    def stop(self, first: int = 0, last: int | None = None) -> dict:
        out = self.summarise(first, last)
        if self.proc is not None:
            self.proc.terminate()
        return out


# ---------------------------------------------------------------------------------------------------------
# model / data builders
# ---------------------------------------------------------------------------------------------------------
def build_model(name: str, device, *, num_classes="cfg", dropout=None, use_uncertainty=None, seed=42):
    import torch
    from tinyedm_b200 import configs
    cfg = {"cifar": configs.CIFAR10, "mnist": configs.MNIST, "imagenet": configs.IMAGENET}[name]
    torch.manual_seed(seed)
    model = configs.build_edm(cfg, num_classes=num_classes, dropout_rate=dropout, use_uncertainty=use_uncertainty)
    with torch.no_grad():
        model.denoiser.gain_out.fill_(1.0)   # init value 0 would make D == c_skip*x and every network gradient exactly zero
        if model.u is not None:
            model.u.gain.fill_(1.0)          # same for the uncertainty head (networks.py:96)
    return model.to(device), cfg


def cifar_edm(T, num_classes, dropout, device, seed=42):
    return build_model("cifar", device, num_classes=num_classes, dropout=dropout, seed=seed)[0]


def synth_batch(cfg, n, gen, n_classes):
    import torch
    C, H, W = cfg["image"]
    x = (0.5 * torch.randn(n, C, H, W, generator=gen)).clamp_(-1, 1).pin_memory()
    y = torch.randint(0, max(n_classes or 1, 1), (n,), generator=gen).pin_memory()
    return x, y


class LaunchCounter:
    """Counts C-ABI calls (== kernel launches of OUR library; multi-kernel entry points weighted)."""
    WEIGHT = {"tedm_attention_forward": 2, "tedm_attention_backward": 4, "tedm_init": 0}

    def __init__(self, lib):
        self.lib, self.n, self.orig = lib, 0, lib.call

    def __enter__(self):
        def counted(name, *a):
            self.n += self.WEIGHT.get(name, 1)
            return self.orig(name, *a)
        self.lib.call = counted
        return self

    def __exit__(self, *exc):
        self.lib.call = self.orig


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own modules (oracle/_ref bytecode of src/tinyedm/networks.py, solvers.py) when built,
# else the oracle port; the step glue that needs lightning/torchmetrics (edm.py:84-93, :205-219; metric.py:8-18) is the
# oracle's restatement in both cases
# ---------------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, name: str, threads: int):
        import torch
        from oracle import edm2_oracle as O
        from oracle import ref_loader
        torch.set_num_threads(threads)
        self.O, self.torch, self.name = O, torch, name
        self.cfg = {"cifar": O.CIFAR10, "mnist": O.MNIST, "imagenet": O.IMAGENET}[name]
        gen = torch.Generator().manual_seed(42)
        self.dp = O.init_denoiser_params(self.cfg["denoiser"], gen, gain_out=1.0)
        self.ep = O.init_embedding_params(self.cfg["embedding"], gen)
        self.gen = gen
        self.kind = "port"
        self.den = self.emb = None
        if ref_loader.load() is not None:
            self.den, self.emb = ref_loader.reference_edm_parts(self.cfg)
            ref_loader.load_params(self.den, self.dp)
            ref_loader.load_params(self.emb, self.ep)
            self.kind = "reference"
        self.opt = None

    def _data(self, batch):
        torch = self.torch
        C, H, W = self.cfg["image"]
        clean = (0.5 * torch.randn(batch, C, H, W, generator=self.gen)).clamp(-1, 1)
        ncls = self.cfg["embedding"].num_classes
        labels = torch.randint(0, ncls, (batch,), generator=self.gen) if ncls else None
        return clean, labels

    def train_step_fn(self, batch: int):
        """One optimiser step of edm.py:205-236 + Adam on `batch` synthetic images (fp32, all host threads)."""
        torch, O, cfg = self.torch, self.O, self.cfg
        clean, labels = self._data(batch)
        sd = cfg["denoiser"].sigma_data
        if self.kind == "reference":
            den, emb = self.den.train(), self.emb.train()
            plist = list(emb.parameters()) + list(den.parameters())
            opt = torch.optim.Adam(plist, lr=0.02 * LR_RAMP, betas=(0.9, 0.999))
            ema = [q.detach().clone() for q in plist]          # EMAOptimizer's shadow copy (ema.py:270-284)

            def step():
                noisy, sigma = O.diffuse(clean, torch.randn(batch), torch.randn_like(clean), cfg["P_mean"], cfg["P_std"])
                _, e = emb(sigma, labels)
                D = den(noisy, sigma, e)                 # training mode: forced weight norm + dropout inside the modules
                loss = O.training_loss(O.loss_weight(sigma, sd), D, clean)
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
                with torch.no_grad():
                    torch._foreach_lerp_(ema, [q.detach() for q in plist], 1.0 - EMA_GAMMA)
                return float(loss)
            return step
        dp, ep = self.dp, self.ep
        params = [v.requires_grad_(True) for v in dp.values()] + [v.requires_grad_(True) for k, v in ep.items() if k.endswith("weight")]
        opt = torch.optim.Adam(params, lr=0.02 * LR_RAMP, betas=(0.9, 0.999))
        ema = [q.detach().clone() for q in params]
        weights = [v for k, v in list(dp.items()) + list(ep.items()) if k.endswith("weight")]
        p_drop = cfg["denoiser"].dropout_rate
        drop = (lambda t: torch.nn.functional.dropout(t, p_drop, True)) if p_drop > 0 else None

        def step():
            for w in weights:                        # training-mode forced weight normalisation (networks.py:32-34)
                O.forced_weight_norm_(w)
            noisy, sigma = O.diffuse(clean, torch.randn(batch), torch.randn_like(clean), cfg["P_mean"], cfg["P_std"])
            _, e = O.embedding_forward(ep, cfg["embedding"], sigma, labels)
            D = O.denoiser_forward(dp, cfg["denoiser"], noisy, sigma, e, drop=drop)
            loss = O.training_loss(O.loss_weight(sigma, sd), D, clean)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            with torch.no_grad():
                torch._foreach_lerp_(ema, [q.detach() for q in params], 1.0 - EMA_GAMMA)
            return float(loss)
        return step

    def nfe_seconds(self, batch: int, heun_steps: int = 2) -> float:
        """Seconds per network evaluation inside the reference's own Heun loop (solvers.py:43-59), eval mode, measured on
        a `heun_steps`-step solve (2*heun_steps-1 evaluations)."""
        torch, O, cfg = self.torch, self.O, self.cfg
        C, H, W = cfg["image"]
        x0 = torch.randn(batch, C, H, W, generator=self.gen)
        ncls = cfg["embedding"].num_classes
        labels = torch.randint(0, ncls, (batch, 1), generator=self.gen) if ncls else None
        n_eval = 2 * heun_steps - 1
        with torch.no_grad():
            if self.kind == "reference":
                from oracle import ref_loader
                den, emb = self.den.eval(), self.emb.eval()

                def model(x, s, lab):                # edm.py:280-286
                    _, e = emb(s, lab)
                    return den(x, s, e)
                solver = ref_loader.load().solvers.DeterministicSolver(num_steps=heun_steps)
                t0 = time.perf_counter()
                solver.solve(model, x0, labels)
                return (time.perf_counter() - t0) / n_eval
            model = lambda x, s, lab: O.edm_forward(self.dp, cfg["denoiser"], self.ep, cfg["embedding"], x, s.reshape(-1), lab)
            t0 = time.perf_counter()
            O.heun_solve(model, x0, labels, num_steps=heun_steps)
            return (time.perf_counter() - t0) / n_eval


def workload_config(world: int, B: int) -> dict:
    """`config` of the headline workload: the SAME dict on both arms (the reference arm times a bounded sample of it and
    says which in `cpu_baseline.sample`); how each arm launches it is reported beside it, not inside."""
    return {"workload": "CIFAR-10 35.6M unconditional EDM2 training step (cifar10.yaml): diffuse+embed+fwd+loss+bwd+"
                        "allreduce+Adam+EMA", "per_gpu_batch": B, "global_batch": B * world, "image": "3x32x32",
            "parallelism": f"dp{world}", "l2": "no explicit flush: each step streams >5 GB of activations (>> 126 MB L2)",
            "weights": "random init, gain_out=1", "dropout": 0.13,
            "lr": f"0.02 x {LR_RAMP} (the ramp-up region of the reference's schedule, edm.py:306-317; no work skipped)"}


def cpu_kind_note(kind: str) -> str:
    return ("the reference's own src/tinyedm/networks.py + solvers.py (bytecode in oracle/_ref) with the Lightning/torchmetrics step "
            "glue restated (oracle)") if kind == "reference" else "oracle port of the reference algorithm (plain torch ops)"


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = 8
    arm = CpuArm("cifar", threads)
    step = arm.train_step_fn(batch)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = batch * args.steps / dt
    sample = (f"CIFAR-10 35.6M EDM2 training step (fwd+bwd+Adam), fp32, batch {batch} per step (per-image cost is flat in batch); "
              + cpu_kind_note(arm.kind))
    # the second half of the metric and the other configs, each on a bounded sample
    t_nfe = arm.nfe_seconds(8, heun_steps=2)
    sampling = {"value": 8 / (63 * t_nfe), "unit": "img/s", "sample": "3 network evaluations of a 2-step Heun solve at batch 8 "
                "(CIFAR net; the class embedding of configs[2] adds a (B,E) row gather), scaled to the 63 evaluations of a 32-step solve", "s_per_nfe": t_nfe}
    extra = {}
    for name, b_train, b_nfe in (("mnist", 8, 0), ("imagenet", 2, 2)):
        try:
            a = CpuArm(name, threads)
            st = a.train_step_fn(b_train)
            st()
            t1 = time.perf_counter(); st(); st(); d = (time.perf_counter() - t1) / 2
            extra[f"{name}_train"] = {"value": b_train / d, "unit": "img/s", "batch": b_train, "steps_timed": 2}
            if b_nfe:
                tn = a.nfe_seconds(b_nfe, heun_steps=1)
                extra[f"{name}_sampling"] = {"value": b_nfe / (63 * tn), "unit": "img/s", "s_per_nfe": tn, "batch": b_nfe}
            del a, st
        except Exception as e:  # noqa: BLE001
            extra[f"{name}_train"] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "img/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(int(os.environ.get("WORLD_SIZE", max(1, args.gpus))), args.batch),
        "sample_of_config": {"where": "host CPU, all threads", "batch_per_step": batch},
        "cpu_baseline": {"value": v, "unit": "img/s", "cores": threads, "kind": arm.kind, "sample": sample},
        "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sampling": sampling, "configs": extra, "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------
class Ctx:
    """Rank / device / timing helpers shared by the legs of one run."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            opts = None
            if os.environ.get("TEDM_NCCL_MAX_CTAS"):     # cap the SMs NCCL takes from the persistent conv kernels (A/B switch)
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = int(os.environ["TEDM_NCCL_MAX_CTAS"])
                opts.config.min_ctas = min(opts.config.max_ctas, 1)
            dist.init_process_group("nccl", device_id=self.dev, pg_options=opts)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        if self.world == 1:
            return ms
        t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def timed(self, fn, steps: int) -> float:
        """ms for `steps` calls: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))


def bench_train_config(cx: Ctx, T, name: str, steps: int, warmup: int, *, use_uncertainty: bool = False) -> dict:
    """Training img/s of one BASELINE config through the package's public step (GraphedTrainStep: graph replay of
    zero_grad + training_step(s) + backward (+ exchange) + one optimiser launch), fed from PINNED HOST batches with the
    loss read back every step (an end-to-end number)."""
    import torch
    from tinyedm_b200.parallel import DistributedEDM
    wl = WORKLOADS[name]
    B, k = wl["batch"], wl["accumulate"]
    model, cfg = build_model(name, cx.dev, use_uncertainty=use_uncertainty or None)
    model.train()
    ddp = DistributedEDM(model) if cx.world > 1 else None
    opt = model.configure_optimizers()["optimizer"]
    base_lr = opt.param_groups[0]["lr"]
    for g in opt.param_groups:
        g["lr"] = base_lr * LR_RAMP
    gen = torch.Generator().manual_seed(4321 + cx.rank)
    ncls = cfg["embedding"]["num_classes"]
    pool = [synth_batch(cfg, B * k, gen, ncls) for _ in range(2)]
    gstep = T.GraphedTrainStep(model, opt, pool[0], ddp=ddp, accumulate=k)
    loss_host = torch.zeros(steps, 1).pin_memory()

    def step(i):
        loss = gstep(pool[i % 2])
        loss_host[i % steps].copy_(loss.detach(), non_blocking=True)
    for i in range(warmup):
        step(i)
    ms = cx.timed(step, steps)
    imgs = cx.world * B * k * steps
    value = imgs / (ms / 1e3)
    pk = peaks()
    peak_tf = pk.get("bf16_tflops_sustained") or 1400.0
    tf = value / cx.world * 3 * wl["fwd_gflop"] / 1e3
    C, H, W = cfg["image"]
    out = {"value": value, "unit": "img/s", "ms_per_step": ms / steps, "per_gpu_batch": B, "accumulate_grad_batches": k,
           "images_per_optimizer_step": cx.world * B * k, "steps": steps, "image": f"{C}x{H}x{W}",
           "params_M": sum(p.numel() for p in model.parameters()) / 1e6, "use_uncertainty": bool(use_uncertainty),
           "tflops_per_gpu": tf, "frac_of_bf16_sustained": tf / peak_tf, "final_loss": float(loss_host[(steps - 1) % steps]),
           "launch": "CUDA graph replay" if gstep.graph is not None else f"eager ({gstep.error})",
           "gpu_launches_per_step": gstep.launches_per_step, "input": "pinned host batch -> device inside the timed region",
           "h2d_bytes_per_step": B * k * (C * H * W * 4 + 8), "d2h_bytes_per_step": 4}
    del gstep, opt, ddp, model
    torch.cuda.empty_cache()
    return out


def bench_sampling_config(cx: Ctx, T, name: str, solves: int) -> dict:
    """32-step Heun img/s (63 network evaluations per solve) of one config, batch-sharded over the ranks with no
    collective: pinned host noise in -> pinned host images out."""
    import torch
    from tinyedm_b200.parallel import shard_slice
    wl = WORKLOADS[name]
    Bs = wl["sample_batch"]
    model, cfg = build_model(name, cx.dev, num_classes=10 if name == "cifar" else "cfg", dropout=0.0, seed=43)
    model.eval()
    solver = T.DeterministicSolver(num_steps=SAMPLE_STEPS)
    C, H, W = cfg["image"]
    ncls = 10 if name == "cifar" else cfg["embedding"]["num_classes"]
    gen = torch.Generator().manual_seed(99)
    x0_all = torch.randn(Bs * cx.world, C, H, W, generator=gen)
    lab_all = torch.randint(0, ncls, (Bs * cx.world, 1), generator=gen)
    sl = shard_slice(Bs * cx.world, cx.rank, cx.world)
    x0_h, lab_h = x0_all[sl].contiguous().pin_memory(), lab_all[sl].contiguous().pin_memory()
    out_h = torch.empty_like(x0_h).pin_memory()

    def sample_once(_i):
        x0 = x0_h.to(cx.dev, non_blocking=True)
        lab = lab_h.to(cx.dev, non_blocking=True)
        out = solver.solve(model, x0, lab)
        out_h.copy_(out, non_blocking=True)
    sample_once(0)                                   # eager warm-up + graph capture + first replay
    ms = cx.timed(sample_once, solves)
    value = cx.world * Bs * solves / (ms / 1e3)
    pk = peaks()
    peak_tf = pk.get("bf16_tflops_sustained") or 1400.0
    tf = value / cx.world * 63 * wl["fwd_gflop"] / 1e3
    res = {"value": value, "unit": "img/s", "workload": f"{name} class-conditional 32-step Heun (63 NFE), batch {Bs}/GPU, "
           "batch-sharded, host x0 in -> host images out", "ms_per_solve": ms / solves, "solves": solves,
           "finite": bool(torch.isfinite(out_h).all()), "tflops_per_gpu": tf, "frac_of_bf16_sustained": tf / peak_tf}
    if name == "imagenet":
        res["minutes_for_50000_images"] = 50000 / value / 60
        res["sample"] = f"{solves} solves of {Bs} images per GPU (the 50 000-image sweep is this rate sustained; no exchange between ranks)"
    del solver, model
    torch.cuda.empty_cache()
    return res


def practical_bar(cx: Ctx) -> dict:
    """The incumbent on the same GPU: the reference's own modules (oracle/_ref bytecode; else the oracle
    port) run eagerly by PyTorch — bf16 autocast + fused Adam for the training step at B = 256, eval forward at B = 128."""
    import torch
    from oracle import edm2_oracle as O
    from oracle import ref_loader
    dev = cx.dev
    if ref_loader.load() is None:
        return {"unavailable": "oracle/_ref not built on this box"}
    torch.backends.cudnn.benchmark = True
    cfg = O.CIFAR10
    gen = torch.Generator().manual_seed(42)
    nets = ref_loader.load().networks
    if not getattr(nets, "_tedm_lerp_shim", False):
        # torch 2.11 refuses Tensor.lerp with mixed dtypes (fp32 `self`, bf16 `end`), which the reference's mp_add
        # (networks.py:87-88) meets under bf16 autocast where one operand comes out of an autocast-off island; torch
        # 2.2.1 (the reference's pin) promoted. One cast, no arithmetic changed.
        orig_mp_add = nets.mp_add
        nets.mp_add = lambda a, b, t=0.5: orig_mp_add(a, b.to(a.dtype), t)
        nets._tedm_lerp_shim = True
    den, emb = ref_loader.reference_edm_parts(cfg)
    ref_loader.load_params(den, O.init_denoiser_params(cfg["denoiser"], gen, gain_out=1.0))
    ref_loader.load_params(emb, O.init_embedding_params(cfg["embedding"], gen))
    den, emb = den.to(dev).train(), emb.to(dev).train()
    opt = torch.optim.Adam(list(den.parameters()) + list(emb.parameters()), lr=0.02 * LR_RAMP, fused=True)
    B = TRAIN_BATCH
    clean = (0.5 * torch.randn(B, 3, 32, 32, device=dev)).clamp(-1, 1)

    def train_step():
        noisy, sigma = O.diffuse(clean, torch.randn(B, device=dev), torch.randn_like(clean), -1.2, 1.2)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            _, e = emb(sigma, None)
            D = den(noisy, sigma, e)
        loss = O.training_loss(O.loss_weight(sigma, 0.5), D.float(), clean)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    def timeit(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n
    t_train = timeit(train_step, 5)
    den.eval(); emb.eval()
    x = torch.randn(SAMPLE_BATCH, 3, 32, 32, device=dev)
    sig = torch.full((SAMPLE_BATCH,), 1.5, device=dev)

    @torch.no_grad()
    def nfe():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            _, e = emb(sig, None)
            return den(x, sig, e)
    t_nfe = timeit(nfe, 5)
    del opt, den, emb
    torch.cuda.empty_cache()
    return {"what": "the reference's own Denoiser/Embedding modules (oracle/_ref) eager on this B200: torch 2.11, bf16 autocast, "
            "cudnn.benchmark, Adam(fused=True)", "train_img_s": B / t_train, "train_ms_per_step": t_train * 1e3,
            "nfe_ms_b128": t_nfe * 1e3, "heun32_img_s": SAMPLE_BATCH / (63 * t_nfe)}


def run_b200(args) -> None:
    import torch
    import torch.distributed as dist
    import tinyedm_b200 as T
    from tinyedm_b200 import _lib, ops
    from tinyedm_b200.parallel import DistributedEDM

    cx = Ctx()
    world, rank, dev = cx.world, cx.rank, cx.dev
    if world != args.gpus and rank == 0:
        print(f"bench.py: WORLD_SIZE={world} but --gpus {args.gpus}; reporting n_gpus={world}", file=sys.stderr)
    ops.ensure_device(dev)
    barrier, timed = cx.barrier, cx.timed

    # ---------------- training (headline: CIFAR-10, cifar10.yaml) ----------------
    B = args.batch
    model = cifar_edm(T, None, 0.13, dev).train()
    ddp = DistributedEDM(model) if world > 1 else None
    opt = model.configure_optimizers()["optimizer"]
    for g in opt.param_groups:
        g["lr"] = 0.02 * LR_RAMP
    n_pool = 4
    gen = torch.Generator().manual_seed(1234 + rank)
    host_imgs = [(0.5 * torch.randn(B, 3, 32, 32, generator=gen)).clamp_(-1, 1).pin_memory() for _ in range(n_pool)]
    host_lbls = [torch.randint(0, 10, (B,), generator=gen).pin_memory() for _ in range(n_pool)]
    dev_imgs = [t.to(dev) for t in host_imgs]
    dev_lbls = [t.to(dev) for t in host_lbls]
    loss_host = torch.zeros(max(args.steps, 1), 1).pin_memory()

    def train_step(batch):
        opt.zero_grad(set_to_none=True)
        loss = model.training_step(batch, 0)
        loss.backward()
        if ddp is not None:
            ddp.finish_backward()
        opt.step()
        return loss

    clocks = ClockSampler(cx.local)
    clocks.start()                      # nvidia-smi needs a few hundred ms to produce its first sample: start early
    for i in range(args.warmup):
        train_step((dev_imgs[i % n_pool], dev_lbls[i % n_pool]))
    torch.cuda.synchronize()
    # the package's graph-replayed step (tinyedm_b200.GraphedTrainStep): same kernels, one graph launch per step
    gstep = None
    if not args.eager:
        gstep = T.GraphedTrainStep(model, opt, (dev_imgs[0], dev_lbls[0]), ddp=ddp)
        if gstep.graph is None:
            if rank == 0:
                print(f"bench.py: CUDA-graph capture unavailable ({gstep.error}); timing the eager step", file=sys.stderr)
            gstep = None
    run_step = gstep if gstep is not None else train_step
    for i in range(args.warmup):
        run_step((dev_imgs[i % n_pool], dev_lbls[i % n_pool]))
    torch.cuda.synchronize()
    t_wait = time.time()
    while clocks.mark() == 0 and time.time() - t_wait < 3.0:     # nvidia-smi needs a moment for its first sample
        time.sleep(0.05)
    # Optional untimed extra steps directly before the timed region (--settle-steps). Measured A/B on one box: 0 vs 64 steps
    # gave 11 754 vs 11 700 img/s — there is no power-cap transient to wait out, so the default is 0.
    settle_steps = max(0, args.settle_steps)      # a COUNT, identical on every rank (each step holds collectives)
    for i in range(settle_steps):
        run_step((dev_imgs[i % n_pool], dev_lbls[i % n_pool]))
    c0 = clocks.mark()
    with LaunchCounter(_lib) as lc:
        ms_dev = timed(lambda i: run_step((dev_imgs[i % n_pool], dev_lbls[i % n_pool])), args.steps)
    # under graph replay the C ABI is not re-entered: count the launches captured in the graph (+ the optimiser's)
    launches = lc.n if gstep is None else gstep.launches_per_step * args.steps
    time.sleep(0.12)                    # let the sample that covers the end of the region arrive
    c1 = clocks.mark()

    def e2e_step(i):
        if gstep is not None:            # pinned host batch -> the step's static device buffers (H2D inside the region)
            loss = gstep((host_imgs[i % n_pool], host_lbls[i % n_pool]))
        else:
            x = host_imgs[i % n_pool].to(dev, non_blocking=True)
            y = host_lbls[i % n_pool].to(dev, non_blocking=True)
            loss = train_step((x, y))
        loss_host[i].copy_(loss.detach(), non_blocking=True)
    ms_e2e = timed(e2e_step, args.steps)
    clk = clocks.summarise(c0, max(c1, c0 + 1))
    final_loss = float(loss_host[args.steps - 1])
    if os.environ.get("TEDM_BENCH_DIAG") == "1" and rank == 0:
        # diagnosis only (stderr): the same two windows again, to tell an ordering effect from a difference of the input paths
        again = [(n, timed(f, args.steps) / args.steps) for n, f in
                 (("device", lambda i: run_step((dev_imgs[i % n_pool], dev_lbls[i % n_pool]))), ("host", e2e_step)) * 2]
        print(f"bench.py diag: ms/step device {ms_dev / args.steps:.3f}, host {ms_e2e / args.steps:.3f}, then "
              + ", ".join(f"{n} {t:.3f}" for n, t in again), file=sys.stderr)
    value = world * B * args.steps / (ms_dev / 1e3)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)

    # ---------------- dominant kernel, timed live with CUDA events on its launch stream ----------------
    eng = model.denoiser.engine
    slot = eng.blocks[0].w["conv_3x3_1"]
    xdom = torch.randn(DOM["B"], DOM["H"], DOM["W"], DOM["Cin"], device=dev).to(torch.bfloat16)
    ydom = torch.empty(DOM["B"], DOM["H"], DOM["W"], DOM["Cout"], device=dev, dtype=torch.bfloat16)
    dom_events = []
    orig_conv = ops.conv2d
    flavour = {0: "plain", 1: "fwd+modulation*silu*dropout", 2: "fwd+mp_add", 3: "dgrad+modsilu adjoint", 4: "dgrad+silu/pixelnorm adjoint"}

    def conv_probe(x, w, ksize, cout, **kw):
        if ksize == 3 and tuple(x.shape) == (DOM["B"], DOM["H"], DOM["W"], DOM["Cin"]) and cout == DOM["Cout"]:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = orig_conv(x, w, ksize, cout, **kw)
            b.record()
            dom_events.append((a, b, kw.get("epi", 0)))
            return out
        return orig_conv(x, w, ksize, cout, **kw)
    import tinyedm_b200.engine as engine_mod
    engine_mod.ops.conv2d = conv_probe
    try:
        if B == DOM["B"]:
            # Instrumented EAGER steps (not part of the reported step time). CUDA events bracket each launch on its
            # stream; to make the interval between them pure kernel time the host has to run ahead of the GPU, so each
            # step is queued behind ~10 ms of ballast GEMMs (otherwise the first forward launches of a step are issued
            # into an empty queue and the interval includes Python launch latency).
            # Two un-instrumented eager steps first: after graph capture the eager allocator pool has to grow again, and
            # a cudaMalloc between two events would be charged to the kernel.
            ballast = torch.empty(8192, 8192, device=dev, dtype=torch.bfloat16).normal_()
            for i in range(2 + min(3, args.steps)):
                if i == 2:
                    torch.cuda.synchronize()
                    dom_events.clear()
                for _ in range(12):
                    torch.mm(ballast, ballast)
                train_step((dev_imgs[i % n_pool], dev_lbls[i % n_pool]))
            del ballast
        else:
            for _ in range(30):
                conv_probe(xdom, slot.fwd, 3, DOM["Cout"], out=ydom)
    finally:
        engine_mod.ops.conv2d = orig_conv
    torch.cuda.synchronize()
    dom_ms = [a.elapsed_time(b) for a, b, _ in dom_events]
    dom_avg = sum(dom_ms) / len(dom_ms)
    pk = peaks()
    peak_tf = pk.get("bf16_tflops_sustained") or 1400.0
    achieved_tf = DOM_FLOP / (dom_avg * 1e-3) / 1e12
    by_epi = {}
    for (a, b, e) in dom_events:
        by_epi.setdefault(e, []).append(a.elapsed_time(b))
    per_flavour = {flavour.get(e, str(e)): {"launches": len(v), "ms_avg": sum(v) / len(v),
                                            "tflops": DOM_FLOP / (sum(v) / len(v) * 1e-3) / 1e12} for e, v in sorted(by_epi.items())}
    traffic = dominant_kernel_traffic()
    roofline = {"bound": "tensor", "kernel": "conv_pair_kernel<*> (tcgen05 cta_group::2 implicit GEMM) 3x3 256->256 @32x32 B256, every "
                "launch of this shape in a training step (forward and data-gradient, with their fused epilogues)",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if pk else "fallback (B200_PROFILING.md)",
                "launch_ms_avg": dom_avg, "launches_timed": len(dom_ms), "traffic": traffic.get("traffic_bytes"),
                "traffic_source": traffic.get("source"), "per_epilogue": per_flavour}

    # ---------------- sampling (configs[2]) and the other BASELINE configs ----------------
    graph_mode = gstep is not None
    del opt, gstep, run_step, ddp, model, eng, slot
    torch.cuda.empty_cache()
    n_solves = max(1, min(3, args.steps))
    sampling = bench_sampling_config(cx, T, "cifar", n_solves)
    extras: dict = {}
    if not args.headline_only:
        k_x = max(3, min(args.steps, 8))
        for key, fn in (
                ("cifar_train_uncertainty", lambda: bench_train_config(cx, T, "cifar", k_x, 3, use_uncertainty=True)),
                ("mnist_train", lambda: bench_train_config(cx, T, "mnist", k_x, 3)),
                ("imagenet_train", lambda: bench_train_config(cx, T, "imagenet", max(3, min(args.steps, 5)), 3)),
                ("imagenet_sampling", lambda: bench_sampling_config(cx, T, "imagenet", 2))):
            try:
                extras[key] = fn()
            except Exception as e:  # noqa: BLE001 - an extra config must never take the headline line down
                extras[key] = {"error": f"{type(e).__name__}: {e}"}
                torch.cuda.synchronize()
                torch.cuda.empty_cache()
    clocks.stop()

    # ---------------- practical bar + CPU baseline (rank 0, N=1 only) ----------------
    bar = cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            bar = practical_bar(cx)
        except Exception as e:  # noqa: BLE001
            bar = {"error": f"{type(e).__name__}: {e}"}
        threads = os.cpu_count() or 1
        arm = CpuArm("cifar", threads)
        step = arm.train_step_fn(8)
        step()
        t0 = time.perf_counter()
        reps = 0
        while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 12):
            step()
            reps += 1
        dt = time.perf_counter() - t0
        t_nfe = arm.nfe_seconds(8, heun_steps=2)
        cpu = {"value": 8 * reps / dt, "unit": "img/s", "cores": threads, "kind": arm.kind,
               "sample": f"{reps} CPU training steps of the same CIFAR config at batch 8 (fp32, all host threads); " + cpu_kind_note(arm.kind),
               "sampling_value": 8 / (63 * t_nfe), "sampling_sample": "3 network evaluations of a 2-step Heun solve at batch 8, scaled to 63"}

    if rank == 0:
        train_tflops = value * 3 * FWD_GFLOP_PER_IMG / 1e3
        line = {
            "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world, B),
            "launch": {"mode": "CUDA graph replay of fwd+bwd (GraphedTrainStep) + 1 fused Adam/EMA launch" if graph_mode
                       else "eager (one C-ABI call per kernel)", "settle_steps": settle_steps},
            "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": B * 3 * 32 * 32 * 4 + B * 8,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "practical_bar": bar,
            "train_tflops_per_gpu": train_tflops / world,
            "train_frac_of_bf16_sustained": train_tflops / world / peak_tf,
            "final_loss": final_loss,
            "sampling": sampling, "configs": extras,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=TRAIN_BATCH, help="per-GPU training batch (the metric is quoted on 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="skip the other BASELINE configs (MNIST / ImageNet-latent / uncertainty)")
    ap.add_argument("--eager", action="store_true", help="time the eager step instead of the CUDA-graph replay")
    ap.add_argument("--settle-steps", type=int, default=0, help="extra untimed steps directly before the timed region")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
