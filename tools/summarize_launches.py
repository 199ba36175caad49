"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share."""
import csv, re, sys, collections
path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iu = hdr.index("Metric Unit")
tot = collections.OrderedDict()
n = 0
for r in rows[1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("tedm::(anonymous namespace)::", "").replace("unnamed>::", "")
    c, t = tot.get(name, (0, 0.0))
    tot[name] = (c + 1, t + us)
    n += 1
total = sum(t for _, t in tot.values())
print(f"# {path}: {n} launches, {total/1e3:.2f} ms of kernel time (ncu-serialised, cold-cache: compare shares, not absolutes)")
print(f"{'kernel':60s} {'n':>5s} {'total us':>10s} {'avg us':>9s} {'share':>7s}")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {c:5d} {t:10.1f} {t/c:9.1f} {100*t/total:6.1f}%")
