#!/usr/bin/env python
"""Headline benchmark of the tinyedm_b200 hot path (contract: see the task's bench.py section / DESIGN.md §measurement).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

Workload (BASELINE.json configs[1], experiments/conf/cifar10.yaml): CIFAR-10 35.6 M unconditional EDM2 training,
per-GPU batch 256, bf16 network / fp32 master weights, dropout 0.13, fused Adam(lr 0.02) + EMA(0.13); a "step" is
noise -> embedding -> denoiser forward -> loss -> backward -> gradient all-reduce (N>1) -> optimiser, on synthetic
images 0.5*N(0,1) clipped to [-1,1] with random-init weights (gain_out set to 1 so the network is not hidden).
The same line also carries the second half of the metric: class-conditional 32-step Heun sampling (configs[2],
batch 128 per GPU, 63 network evaluations), batch-sharded over the ranks.
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
FWD_GFLOP_PER_IMG = 27.001          # BASELINE.md §2 (conv3x3 24.707 + conv1x1 1.931 + attention 0.361 + linear)
# dominant kernel: conv_pair_kernel, 3x3 256->256 at 32x32, per-GPU batch 256
DOM = dict(B=TRAIN_BATCH, H=32, W=32, Cin=256, Cout=256, k=3)
DOM_FLOP = 2.0 * DOM["B"] * DOM["H"] * DOM["W"] * DOM["Cin"] * DOM["Cout"] * DOM["k"] ** 2


def peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        return {}


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

    def stop(self, first: int = 0, last: int | None = None) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
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


# ---------------------------------------------------------------------------------------------------------
# model / data builders
# ---------------------------------------------------------------------------------------------------------
def cifar_edm(T, num_classes, dropout, device, seed=42):
    import torch
    from tinyedm_b200.configs import CIFAR10, build_edm
    torch.manual_seed(seed)
    model = build_edm(CIFAR10, num_classes=num_classes, dropout_rate=dropout)
    with torch.no_grad():
        model.denoiser.gain_out.fill_(1.0)   # init value 0 would make D == c_skip*x and every network gradient exactly zero
    return model.to(device)


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
# CPU baseline (oracle port == the reference's algorithm in plain torch ops on the host cores)
# ---------------------------------------------------------------------------------------------------------
def cpu_train_step_fn(batch: int, threads: int):
    import torch
    from oracle import edm2_oracle as O
    torch.set_num_threads(threads)
    cfg = O.CIFAR10
    gen = torch.Generator().manual_seed(42)
    dp = O.init_denoiser_params(cfg["denoiser"], gen, gain_out=1.0)
    ep = O.init_embedding_params(cfg["embedding"], gen)
    params = [v.requires_grad_(True) for k, v in dp.items()] + [ep["sigma_embed.weight"].requires_grad_(True)]
    opt = torch.optim.Adam(params, lr=0.02, betas=(0.9, 0.999))
    weights = [v for k, v in list(dp.items()) + list(ep.items()) if k.endswith("weight")]
    clean = (0.5 * torch.randn(batch, 3, 32, 32, generator=gen)).clamp(-1, 1)

    def step():
        for w in weights:                        # training-mode forced weight normalisation (networks.py:32-34)
            O.forced_weight_norm_(w)
        noisy, sigma = O.diffuse(clean, torch.randn(batch), torch.randn_like(clean), -1.2, 1.2)
        _, e = O.embedding_forward(ep, cfg["embedding"], sigma)
        drop = lambda t: torch.nn.functional.dropout(t, 0.13, True)
        D = O.denoiser_forward(dp, cfg["denoiser"], noisy, sigma, e, drop=drop)
        loss = O.training_loss(O.loss_weight(sigma, 0.5), D, clean)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return float(loss)
    return step


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = 8
    step = cpu_train_step_fn(batch, threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = batch * args.steps / dt
    sample = f"CIFAR-10 35.6M EDM2 training step (fwd+bwd+Adam), fp32, batch {batch} per step (per-image cost is flat in batch)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "img/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CIFAR-10 35.6M unconditional EDM2 training (cifar10.yaml), CPU oracle port", "batch_per_step": batch},
        "cpu_baseline": {"value": v, "unit": "img/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist
    import tinyedm_b200 as T
    from tinyedm_b200 import _lib, ops
    from tinyedm_b200.parallel import DistributedEDM, shard_slice

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"bench.py: WORLD_SIZE={world} but --gpus {args.gpus}; reporting n_gpus={world}", file=sys.stderr)
    ops.ensure_device(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---------------- training ----------------
    B = args.batch
    model = cifar_edm(T, None, 0.13, dev).train()
    ddp = DistributedEDM(model) if world > 1 else None
    opt = model.configure_optimizers()["optimizer"]
    for g in opt.param_groups:
        g["lr"] = 0.02 * 1e-3          # a ramp-up learning rate: keeps random-init training numerically tame over the run
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

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    clocks = ClockSampler(local)
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
    while clocks.mark() == 0 and time.time() - t_wait < 3.0:
        time.sleep(0.05)
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
    clk = clocks.stop(c0, max(c1, c0 + 1))
    final_loss = float(loss_host[args.steps - 1])
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
    roofline = {"bound": "tensor", "kernel": "conv_pair_kernel<*> (tcgen05 cta_group::2 implicit GEMM) 3x3 256->256 @32x32 B256, every "
                "launch of this shape in a training step (forward and data-gradient, with their fused epilogues)",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if pk else "fallback (B200_PROFILING.md)",
                "launch_ms_avg": dom_avg, "launches_timed": len(dom_ms), "traffic": TRAFFIC_BYTES,
                "per_epilogue": per_flavour}

    # ---------------- sampling (configs[2]) ----------------
    graph_mode = gstep is not None
    del opt, gstep, run_step
    model.eval()
    smodel = cifar_edm(T, 10, 0.0, dev).eval()
    solver = T.DeterministicSolver(num_steps=SAMPLE_STEPS)
    gen2 = torch.Generator().manual_seed(99)
    x0_all = torch.randn(SAMPLE_BATCH * world, 3, 32, 32, generator=gen2)
    lab_all = torch.randint(0, 10, (SAMPLE_BATCH * world, 1), generator=gen2)
    sl = shard_slice(SAMPLE_BATCH * world, rank, world)
    x0_h, lab_h = x0_all[sl].contiguous().pin_memory(), lab_all[sl].contiguous().pin_memory()
    out_h = torch.empty_like(x0_h).pin_memory()

    def sample_once(_i):
        x0 = x0_h.to(dev, non_blocking=True)
        lab = lab_h.to(dev, non_blocking=True)
        out = solver.solve(smodel, x0, lab)
        out_h.copy_(out, non_blocking=True)
    sample_once(0)
    n_solves = max(1, min(3, args.steps))
    ms_s = timed(sample_once, n_solves)
    sample_value = world * SAMPLE_BATCH * n_solves / (ms_s / 1e3)
    finite = bool(torch.isfinite(out_h).all())

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        step = cpu_train_step_fn(8, threads)
        step()
        t0 = time.perf_counter()
        reps = 0
        while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 12):
            step()
            reps += 1
        dt = time.perf_counter() - t0
        cpu = {"value": 8 * reps / dt, "unit": "img/s", "cores": threads, "kind": "port",
               "sample": f"{reps} CPU training steps of the same CIFAR config at batch 8 (fp32 oracle port, all host threads)"}

    if rank == 0:
        train_tflops = value * 3 * FWD_GFLOP_PER_IMG / 1e3
        line = {
            "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "CIFAR-10 35.6M unconditional EDM2 training step (cifar10.yaml): diffuse+embed+fwd+loss+bwd+"
                                   "allreduce+fused Adam/EMA", "per_gpu_batch": B, "global_batch": B * world, "image": "3x32x32",
                       "parallelism": f"dp{world}", "l2": "no explicit flush: each step streams >5 GB of activations (>> 126 MB L2)",
                       "weights": "random init, gain_out=1", "dropout": 0.13,
                       "launch": "CUDA graph replay of fwd+bwd (GraphedTrainStep) + 1 optimiser launch" if graph_mode
                                 else "eager (one C-ABI call per kernel)"},
            "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": B * 3 * 32 * 32 * 4 + B * 8,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
            "train_tflops_per_gpu": train_tflops / world,
            "train_frac_of_bf16_sustained": train_tflops / world / peak_tf,
            "final_loss": final_loss,
            "sampling": {"value": sample_value, "unit": "img/s", "workload": "CIFAR-10 class-conditional 32-step Heun (63 NFE), "
                         f"batch {SAMPLE_BATCH}/GPU, batch-sharded, host x0 in -> host images out", "ms_per_solve": ms_s / n_solves,
                         "solves": n_solves, "finite": finite,
                         "tflops_per_gpu": sample_value / world * 63 * FWD_GFLOP_PER_IMG / 1e3},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (conv_pair_kernel, 3x3 256->256 @32x32,
# B=256) from the committed `ncu --set full` capture profiles/r1k_conv_pair_ncu_full.md: launch-weighted mean over the
# four flavours of a training step (347 / 377 / 375 / 517 MB at 9 : 18 : 18 : 9); algorithmic bytes 402-537 MB per launch.
TRAFFIC_BYTES = 394.8e6


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=TRAIN_BATCH, help="per-GPU training batch (the metric is quoted on 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="time the eager step instead of the CUDA-graph replay")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
