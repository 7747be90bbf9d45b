"""Condense an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / bench.py quote.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
print(f"# ncu -i {rep} --page raw   (one launch, --set full --clock-control none)")
for h, u, v in zip(hdr, units, vals):
    if h in want or any(h.startswith(w) and h == w for w in want):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
try:
    h2 = srows[1]
    ia, isrc, iall, iex = h2.index("Address"), h2.index("Source"), h2.index("Warp Stall Sampling (All Samples)"), h2.index("Instructions Executed")
    data = []
    for r in srows[2:]:
        try:
            data.append((int(r[iall]), r[isrc][:100], r[iex]))
        except Exception:
            pass
    tot = sum(d[0] for d in data)
    print(f"# top stall-sample instructions (of {tot} samples); UTCHMMA / UTMALDG / LDTM = tcgen05.mma / TMA / tcgen05.ld")
    for s, text, ex in sorted(data, reverse=True)[:12]:
        print(f"{100 * s / tot:5.1f}%  ex={ex:>9}  {text}")
    ops = {}
    for s, text, ex in data:
        for key in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "HMMA", "STG", "LDG"):
            if key in text:
                ops[key] = ops.get(key, 0) + int(ex or 0)
    print("# executed instruction counts:", ops)
except Exception as e:
    print("# source page unavailable:", e)
