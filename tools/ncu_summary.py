"""Key metrics + hottest SASS lines of every kernel in an .ncu-rep (reads `ncu -i ... --page raw/source --csv`)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__cluster_size", "launch__occupancy_cluster_pct", "launch__occupancy_limit",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__cluster_max_active", "launch__waves_per_multiprocessor"]
idx = [i for i, h in enumerate(hdr) if any(h == w or (w.endswith("limit") and h.startswith(w)) for w in want)]
for r in rows[2:]:
    print("-" * 100)
    for i in idx:
        print(f"  {hdr[i]:80s} {r[i][:70]:>20s} {units[i]}")
if len(sys.argv) > 2:
    k = int(sys.argv[2])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[hi]
    ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = [(int(r[isamp]), r[ia], r[iex]) for r in rows[hi + 1:] if len(r) > isamp and r[isamp].isdigit()]
    half = len(data) // 2 if len(data) > 1 and data[0][1] == data[len(data) // 2][1] else len(data)
    data = data[:half]
    tot = sum(d[0] for d in data)
    print(f"\nlaunch {k}: {tot} samples over {len(data)} SASS instructions; lines with >0.7% or sync/TMA/MMA:")
    for i, (s, a, ex) in enumerate(data):
        if s > tot * 0.007 or any(t in a for t in ("SYNCS", "UTMA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "BAR.", "UBLKCP", "UTMASTG", "DEPBAR")):
            print(f"  {i:5d} {s:7d} {100*s/tot:5.1f}% x{ex:>9s}  {a[:110]}")
