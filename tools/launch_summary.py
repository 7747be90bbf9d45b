"""Condense an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals and shares.
usage: python tools/launch_summary.py gpurun_out/x.csv "<header comment>" > profiles/x_summary.txt"""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
agg = {}
for r in rows:
    name = re.sub(r"\(.*", "", r[4])
    ns = float(r[14].replace(",", ""))
    if r[13] == "us":
        ns *= 1e3
    a = agg.setdefault(name, [0.0, 0])
    a[0] += ns; a[1] += 1
tot = sum(a[0] for a in agg.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
for name, (ns, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{ns / 1e6:8.3f} ms {100 * ns / tot:5.1f}%  x{n:3d}  {name}")
print(f"total {tot / 1e6:.3f} ms over {len(rows)} launches")
