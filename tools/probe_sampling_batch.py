"""GPU probe: ImageNet-latent 32-step Heun sampling throughput of one GPU against the per-solve batch (the 50 000-image sweep of
BASELINE.json configs[4] leaves it free). Prints batch, img/s, ms per solve, fraction of the sustained bf16 peak."""
import sys, json, time
sys.path.insert(0, '.')
import torch, bench
import tinyedm_b200 as T
cx = bench.Ctx()
for bs in (64, 96, 128, 176):
    bench.WORKLOADS["imagenet"]["sample_batch"] = bs
    r = bench.bench_sampling_config(cx, T, "imagenet", 2)
    print(bs, round(r["value"], 2), round(r["ms_per_solve"], 1), round(r["frac_of_bf16_sustained"], 3), flush=True)
