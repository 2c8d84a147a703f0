"""Key metrics of an `ncu --set full` report as a markdown table:  python scripts/ncu_summary.py file.ncu-rep [...]"""
import csv, io, subprocess, sys
WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("lts__t_bytes.sum", "L2 bytes"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
print("| kernel | " + " | ".join(n for _, n in WANT) + " |\n|---|" + "---|" * len(WANT))
for path in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("dbmm::", "")
        if name in seen:
            continue
        seen.add(name)
        cells = []
        for m, _ in WANT:
            if m in ix:
                v, u = r[ix[m]], units[ix[m]]
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:.1f}" if abs(f) < 1e4 else f"{f:.3g}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            else:
                cells.append("-")
        print(f"| `{name[:48]}` | " + " | ".join(cells) + " |")
