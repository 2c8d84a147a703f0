"""Turn the scratch captures of scripts/profile_round.sh (gpurun_out/<tag>_*) into the tracked evidence under profiles/:
  profiles/<tag>_launches_step.csv / .md    ncu launch list of one training step (cold caches, serialised launches)
  profiles/<tag>_ncu_full_summary.md         key `--set full` metrics of every step kernel
  profiles/r1_traffic.json                   (updated) DRAM bytes per launch, read by bench.py's roofline.traffic
  profiles/<tag>_bench.json                  the bench line of the same run
usage: python scripts/summarize_profiles.py <tag>"""
import collections, csv, json, os, shutil, subprocess, sys

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
MULT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}


def short(name):
    return name.replace("void ", "").replace("dbmm::", "").split("(")[0]


# ---- launch list
src = os.path.join(go, f"{tag}_launches_step.csv")
lines = [l for l in open(src) if not l.startswith("==")]
open(os.path.join(pr, f"{tag}_launches_step.csv"), "w").writelines(lines)
rows = list(csv.DictReader(lines))
per = collections.OrderedDict()
for r in rows:
    k = (short(r["Kernel Name"]), r["Grid Size"], r["Block Size"])
    per.setdefault(k, collections.defaultdict(list))[r["Metric Name"]].append(float(r["Metric Value"].replace(",", "")) * MULT.get(r["Metric Unit"], 1))
tot = sum(sum(m["gpu__time_duration.sum"]) / len(m["gpu__time_duration.sum"]) for m in per.values())
md = [f"# ncu launch list of one training step ({tag})", "",
      "`N=20480 EPOCHS=1 DBMM_GRAPH=0 DBMM_TAIL=serial ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
      "--clock-control none -s 300 -c 60 --csv python scripts/train_only.py` on one B200.  Default cache control: every launch starts with "
      "cold caches and the launches are serialised, so compare SHARES with `bench.py`'s `roofline.kernel_us`, not absolutes; "
      "`DBMM_TAIL=serial` puts `k_tail_w2` in line (in the product's epoch graph it runs on a second branch, off the critical path).", "",
      "| kernel | grid | block | launches | mean us (cold) | share | DRAM read / launch | DRAM write / launch |", "|---|---|---|---|---|---|---|---|"]
for (name, grid, block), m in per.items():
    d = sum(m["gpu__time_duration.sum"]) / len(m["gpu__time_duration.sum"])
    rd = sum(m["dram__bytes_read.sum"]) / len(m["dram__bytes_read.sum"]); wr = sum(m["dram__bytes_write.sum"]) / len(m["dram__bytes_write.sum"])
    md.append(f"| `{name}` | {grid} | {block} | {len(m['gpu__time_duration.sum'])} | {d:.2f} | {d / tot:.3f} | {rd / 1e6:.2f} MB | {wr / 1e6:.2f} MB |")
open(os.path.join(pr, f"{tag}_launches_step.md"), "w").write("\n".join(md) + "\n")

# ---- traffic json (kept file name: bench.py reads it)
tj = os.path.join(pr, "r1_traffic.json")
traffic = json.load(open(tj)) if os.path.exists(tj) else {}
for (name, grid, block), m in per.items():
    base = name.split("<")[0]
    rd = sum(m["dram__bytes_read.sum"]) / len(m["dram__bytes_read.sum"]); wr = sum(m["dram__bytes_write.sum"]) / len(m["dram__bytes_write.sum"])
    traffic[base] = round(rd + wr)
    traffic[base + ":detail"] = {"launches": len(m["dram__bytes_read.sum"]), "dram_read_bytes": rd, "dram_write_bytes": wr,
                                 "ncu_duration_us": sum(m["gpu__time_duration.sum"]) / len(m["gpu__time_duration.sum"]), "capture": tag}
json.dump(traffic, open(tj, "w"), indent=1)

# ---- --set full summary
rep = os.path.join(go, f"{tag}_step_full.ncu-rep")
want = collections.OrderedDict([
    ("gpu__time_duration.sum", "duration"), ("sm__cycles_active.avg", "SM active cycles, avg"),
    ("launch__registers_per_thread", "regs/thread"), ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("smsp__inst_executed.sum", "warp instructions")])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(out.splitlines()))
hdr, units = rr[0], rr[1]
ki = hdr.index("Kernel Name")
cols = [(hdr.index(k), v) for k, v in want.items() if k in hdr]
kern = [(short(r[ki]), r) for r in rr[2:]]
md = [f"# ncu --set full, one launch of every step kernel ({tag})", "",
      "`ncu --set full --import-source on --clock-control none -k regex:... -s 120 -c 6 python scripts/train_only.py` "
      f"(stream launches, `DBMM_TAIL=serial`); report: `gpurun_out/{tag}_step_full.ncu-rep` (scratch).  B = 1024, D = 1024, H = 128.", "",
      "| metric | " + " | ".join(f"`{k}`" for k, _ in kern) + " |", "|---|" + "---|" * len(kern)]
for ci, label in cols:
    md.append(f"| {label} ({units[ci]}) | " + " | ".join(r[ci] for _, r in kern) + " |")
open(os.path.join(pr, f"{tag}_ncu_full_summary.md"), "w").write("\n".join(md) + "\n")
shutil.copy(os.path.join(go, f"{tag}_bench.json"), os.path.join(pr, f"{tag}_bench.json"))
print("\n".join(md[-14:]))
