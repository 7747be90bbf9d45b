"""Condense an ncu launch list taken with
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv
over ONE guided step (tools/profile_step.py) into per-kernel DRAM traffic, and write the JSON bench.py quotes as
`roofline.traffic` (sum over the implicit-GEMM launches of the step).
usage: python tools/dram_summary.py gpurun_out/x.csv gpurun_out/algo_bytes.json profiles/r2_igemm_dram_step_b64.json > profiles/x_summary.txt"""
import csv, json, re, sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14]
per = {}          # launch id -> {name, metrics}
for r in rows:
    if r[12] not in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
        continue
    v = float(r[14].replace(",", ""))
    unit = r[13]
    scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d = per.setdefault(r[0], {"name": re.sub(r"\(.*", "", r[4])})
    d[r[12]] = v * scale
agg = {}
for d in per.values():
    a = agg.setdefault(d["name"], [0.0, 0.0, 0.0, 0])
    a[0] += d.get("gpu__time_duration.sum", 0.0)
    a[1] += d.get("dram__bytes_read.sum", 0.0)
    a[2] += d.get("dram__bytes_write.sum", 0.0)
    a[3] += 1
tot_ns = sum(a[0] for a in agg.values())
print("# per-kernel totals of one guided step (ncu, cold-cache serialised launches): time, DRAM read, DRAM write")
for name, (ns, rd, wr, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{ns / 1e6:8.3f} ms {100 * ns / tot_ns:5.1f}%  x{n:3d}  read {rd / 1e6:9.1f} MB  write {wr / 1e6:9.1f} MB  {name}")
all_rd, all_wr = sum(a[1] for a in agg.values()), sum(a[2] for a in agg.values())
print(f"total {tot_ns / 1e6:.3f} ms over {len(per)} launches; DRAM read {all_rd / 1e9:.2f} GB, write {all_wr / 1e9:.2f} GB")
ig = [(k, a) for k, a in agg.items() if "igemm" in k]
out = {"dram_bytes_per_step": int(sum(a[1] + a[2] for _, a in ig)), "launches": int(sum(a[3] for _, a in ig)),
       "kernel_ms_ncu": sum(a[0] for _, a in ig) / 1e6, "whole_step_dram_bytes": int(all_rd + all_wr), "whole_step_launches": len(per),
       "source": sys.argv[1]}
try:
    out.update({k: v for k, v in json.load(open(sys.argv[2])).items() if k in ("algorithmic_bytes_per_step", "tflop_per_step")})
except Exception:
    pass
json.dump(out, open(sys.argv[3], "w"), indent=1)
print("# implicit-GEMM launches:", json.dumps(out))
