"""Per-launch table from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,... --csv` over tools/bench_elementwise.py (EW_ITERS=1):
DRAM traffic, duration, achieved DRAM GB/s and ncu's own percentage of the DRAM peak for every HBM-bound kernel."""
import csv, re, sys, collections, json, os
path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
hdr = rows[0]
iid, ik, ig = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Grid Size")
im, iu, iv = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
L = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv:
        continue
    d = L.setdefault(r[iid], {"name": re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("unnamed>::", ""), "grid": r[ig]})
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    if r[im] == "gpu__time_duration.sum":
        v = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    if r[im].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    d[r[im]] = v
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6551.7
print(f"# {path}: one cold launch per case of tools/bench_elementwise.py (B = 256, CIFAR shapes), ncu --clock-control none")
print(f"# GB/s = (dram read + write) / duration; '% meas' = GB/s / {peak:.0f} (MEASURED_PEAKS.json); '% ncu' = gpu__dram_throughput pct of peak")
print(f"{'kernel':44s} {'grid':>14s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>8s} {'% meas':>7s} {'% ncu':>6s} {'warps%':>7s}")
for d in L.values():
    t = d.get("gpu__time_duration.sum", 0.0)
    rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    gbs = (rd + wr) / t / 1e3 if t else 0.0
    print(f"{d['name'][:44]:44s} {d['grid']:>14s} {t:8.1f} {rd/1e6:8.1f} {wr/1e6:8.1f} {gbs:8.1f} {100*gbs/peak:6.1f}% "
          f"{d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0):5.1f}% {d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):6.1f}%")
