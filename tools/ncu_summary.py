"""one markdown table per .ncu-rep: the metrics DESIGN.md / profiles/*.md quote (read with `ncu -i … --page raw --csv`)"""
import csv, subprocess, sys, io

WANT = [
    ("gpu__time_duration.sum", "time us"),
    ("dram__bytes_read.sum", "DRAM rd MB"),
    ("dram__bytes_write.sum", "DRAM wr MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__inst_executed.sum", "warp inst M"),
    ("sm__instruction_throughput.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem B"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("smsp__cycles_active.avg", "SMSP cyc avg"),
    ("smsp__cycles_active.max", "SMSP cyc max"),
]

def to_num(v, unit, name):
    try:
        x = float(v.replace(",", ""))
    except Exception:
        return v
    u = unit.lower()
    if name.startswith("gpu__time_duration"):
        x = x / 1e3 if u == "ns" else x * 1e3 if u == "ms" else x * 1e6 if u == "s" else x
    elif "bytes" in name and "DRAM" in dict(WANT).get(name, ""):
        x = x * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1e-6)
    elif name == "sm__inst_executed.sum":
        x = x / 1e6
    return x

def main(path, limit=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {}
    for i, h in enumerate(hdr):
        base = h.split(".TriageCompute.")[-1]
        col.setdefault(base, i)
        col.setdefault(h, i)
    names = [n for n, _ in WANT if n in col]
    print("| # | kernel | grid | block | " + " | ".join(dict(WANT)[n] for n in names) + " |")
    print("|---|---|---|---|" + "---|" * len(names))
    for r in data[:limit]:
        k = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for n in names:
            v = to_num(r[col[n]], units[col[n]], n)
            cells.append(f"{v:.4g}" if isinstance(v, float) else str(v))
        print(f"| {r[col['ID']]} | {k} | {r[col['Grid Size']]} | {r[col['Block Size']]} | " + " | ".join(cells) + " |")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
