"""N>1 timeline of one data-parallel CIFAR training step (torchrun, one rank per GPU): CUPTI kernel records of rank 0 for a
few graph-replayed steps, split into this library's kernels (compute stream) and NCCL's, with the EXPOSED communication
time = intervals in which an NCCL kernel runs and no compute kernel does. Writes gpurun_out/ddp_timeline_n{N}.txt.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/profile_ddp.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
import tinyedm_b200 as T
from tinyedm_b200.configs import CIFAR10, build_edm
from tinyedm_b200.parallel import DistributedEDM

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
opts = None
if world > 1:
    if os.environ.get("TEDM_NCCL_MAX_CTAS"):
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = int(os.environ["TEDM_NCCL_MAX_CTAS"]); opts.config.min_ctas = 1
    dist.init_process_group("nccl", device_id=dev, pg_options=opts)
B = 256
torch.manual_seed(0)
model = build_edm(CIFAR10).to(dev).train()
with torch.no_grad():
    model.denoiser.gain_out.fill_(1.0)
ddp = DistributedEDM(model) if world > 1 else None       # TEDM_DDP_BUCKET_MB / TEDM_DDP_INLINE are read by the wrapper
opt = model.configure_optimizers()["optimizer"]
for g in opt.param_groups: g["lr"] = 2e-5
x = (0.5 * torch.randn(B, 3, 32, 32, device=dev)).clamp(-1, 1)
y = torch.zeros(B, dtype=torch.long, device=dev)
step = T.GraphedTrainStep(model, opt, (x, y), ddp=ddp)
assert step.graph is not None, step.error
for _ in range(5): step((x, y))
torch.cuda.synchronize()
if world > 1: dist.barrier()
N_STEPS = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N_STEPS + 2): step((x, y))
    torch.cuda.synchronize()
if rank == 0:
    evs = sorted((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                 if e.device_time_total > 0 and e.time_range.end > e.time_range.start)
    # the ranks enter the profiled region at different moments: the first two steps absorb that skew (their collectives
    # wait for the slowest rank) and are dropped; a step ends with the optimiser kernel
    ends = [b for a, b, n in evs if "adam" in n]
    if len(ends) >= N_STEPS + 2:
        t_cut = ends[1]
        evs = [e for e in evs if e[0] >= t_cut]
    is_nccl = lambda n: "nccl" in n.lower()
    comp = [(a, b, n) for a, b, n in evs if not is_nccl(n)]
    comm = [(a, b, n) for a, b, n in evs if is_nccl(n)]
    span = evs[-1][1] - evs[0][0]
    def union(iv):
        tot, cur_a, cur_b = 0.0, None, None
        for a, b in sorted(iv):
            if cur_b is None or a > cur_b:
                if cur_b is not None: tot += cur_b - cur_a
                cur_a, cur_b = a, b
            else:
                cur_b = max(cur_b, b)
        if cur_b is not None: tot += cur_b - cur_a
        return tot
    busy_comp = union([(a, b) for a, b, _ in comp])
    busy_comm = union([(a, b) for a, b, _ in comm])
    busy_any = union([(a, b) for a, b, _ in evs])
    exposed = busy_any - busy_comp                    # only NCCL running
    idle = span - busy_any
    os.makedirs("gpurun_out", exist_ok=True)
    path = f"gpurun_out/ddp_timeline_n{world}.txt"
    with open(path, "w") as f:
        P = lambda *a: (print(*a), f.write(" ".join(str(t) for t in a) + "\n"))
        P(f"# tools/profile_ddp.py, world {world}, rank 0, {N_STEPS} graph-replayed CIFAR training steps (B=256/GPU), CUPTI kernel records;"
          f" NCCL max_ctas={os.environ.get('TEDM_NCCL_MAX_CTAS', 'default')}, bucket {os.environ.get('TEDM_DDP_BUCKET_MB', '25')} MiB, inline={os.environ.get('TEDM_DDP_INLINE', '0')}")
        P(f"per step: span {span / N_STEPS / 1e3:.3f} ms | compute kernels busy {busy_comp / N_STEPS / 1e3:.3f} ms | NCCL kernels busy "
          f"{busy_comm / N_STEPS / 1e3:.3f} ms ({len(comm) // N_STEPS} kernels) | EXPOSED communication (NCCL running, no compute kernel) "
          f"{exposed / N_STEPS:.1f} us | idle (nothing running) {idle / N_STEPS:.1f} us")
        if comm:
            P("NCCL kernels of the last step (start relative to the step's first kernel, duration, overlap with compute):")
            last = [c for c in comm][-(len(comm) // N_STEPS):]
            # first kernel of the last step: the first compute kernel after the previous optimiser kernel
            t0 = last[0][0]
            for a, b, n in comp:
                if "adam" in n and b < last[0][0]:
                    t0 = b
            for a, b, n in last:
                ov = union([(max(a, ca), min(b, cb)) for ca, cb, _ in comp if cb > a and ca < b])
                P(f"  +{(a - t0) / 1e3:7.3f} ms  {(b - a):8.1f} us  overlapped {ov:8.1f} us  {n[:90]}")
        # the compute kernels that run concurrently with NCCL: do they slow down? compare with their mean duration elsewhere
        by = {}
        for a, b, n in comp:
            conc = any(cb > a and ca < b for ca, cb, _ in comm)
            by.setdefault(n[:70], [[], []])[1 if conc else 0].append(b - a)
        P("compute kernels: mean duration alone vs while an NCCL kernel is running (us)")
        for n, (alone, conc) in sorted(by.items(), key=lambda kv: -sum(kv[1][0]) - sum(kv[1][1]))[:12]:
            if alone and conc:
                P(f"  {sum(alone) / len(alone):8.1f} (n={len(alone):3d})  {sum(conc) / len(conc):8.1f} (n={len(conc):3d})  {n}")
# no barrier / destroy_process_group here: tearing the process group down while a captured graph still references its
# communicator hung the first version of this tool for the whole gpurun limit; every rank just leaves
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
