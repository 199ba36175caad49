"""DRAM traffic per launch of the dominant kernel from an `ncu --set full` capture of tools/ncu_pair.py (run HERE, no GPU):

    python tools/ncu_traffic.py gpurun_out/r2_conv_pair.ncu-rep [commit]   ->  profiles/conv_pair_traffic.json

bench.py cites that file in `roofline.traffic` (the number is measured under ncu, once per kernel change, never inside a
bench run). tools/ncu_pair.py launches, in this order, the epilogues PLAIN, AXPBY, MODSILU (+raw copy, dropout),
MODSILU_BWD, SILU_BWD and SILU_BWD + fused pixel-norm adjoint of conv_pair_kernel on 3x3 256->256 @32x32, B = 256; a
training step launches them 0 : 18 : 9 : 18 : 3 : 6 times on that shape (bench.py `per_epilogue`), which weights the mean."""
import csv, io, json, os, subprocess, sys

rep = sys.argv[1]
commit = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def num(r, name):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "%": 1}.get(u, 1)
    return v * scale
names = ["plain", "fwd+mp_add", "fwd+modulation*silu*dropout (+raw)", "dgrad+modsilu adjoint", "dgrad+silu adjoint", "dgrad+silu/pixelnorm adjoint"]
weights = [0, 18, 9, 18, 3, 6]
launches = []
for r in rows[2:]:
    if "conv_pair_kernel" not in r[col["Kernel Name"]]:
        continue
    launches.append({"kernel": r[col["Kernel Name"]].split("(")[0][-40:],
                     "dram_read_bytes": num(r, "dram__bytes_read.sum"), "dram_write_bytes": num(r, "dram__bytes_write.sum"),
                     "duration_us": num(r, "gpu__time_duration.sum"),
                     "tensor_pipe_active_pct": num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")})
launches = launches[:len(names)]
for l, n in zip(launches, names):
    l["epilogue"] = n
    l["traffic_bytes"] = l["dram_read_bytes"] + l["dram_write_bytes"]
wsum = sum(w for w, _ in zip(weights, launches))
mean = sum(w * l["traffic_bytes"] for w, l in zip(weights, launches)) / max(wsum, 1)
B, HW, C = 256, 1024, 256
act = B * HW * C * 2
algorithmic = {"fwd+mp_add": 3 * act, "fwd+modulation*silu*dropout (+raw)": 3 * act, "dgrad+modsilu adjoint": 3 * act,
               "dgrad+silu adjoint": 4 * act, "dgrad+silu/pixelnorm adjoint": 4 * act, "plain": 2 * act}
out = {"traffic_bytes": mean, "source": f"ncu --set full --clock-control none, tools/ncu_pair.py, {os.path.basename(rep)} at commit {commit}: "
       "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean weighted by the launches of a training step (18:9:18:3:6)",
       "algorithmic_bytes_per_launch": {k: v + 9 * C * C * 2 for k, v in algorithmic.items()}, "launches": launches}
os.makedirs("profiles", exist_ok=True)
json.dump(out, open("profiles/conv_pair_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
