"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list per kernel."""
import csv, sys, re, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr): continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    key = (name, r[ix["Grid Size"]], r[ix["Block Size"]])
    m = r[ix["Metric Name"]]; v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
    a = agg.setdefault(key, collections.defaultdict(list))
    if m == "gpu__time_duration.sum": v = v / 1e3 if u == "ns" else (v if u == "us" else v * 1e3)
    else: v = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1) * v / 1e6
    a[m].append(v)
tot = sum(sum(a["gpu__time_duration.sum"]) for a in agg.values())
print("| kernel | grid | block | launches | mean us | share | DRAM rd MB | DRAM wr MB |\n|---|---|---|---|---|---|---|---|")
for (n, g, b), a in agg.items():
    t = a["gpu__time_duration.sum"]
    print(f"| `{n[:60]}` | {g} | {b} | {len(t)} | {sum(t)/len(t):.2f} | {sum(t)/tot:.3f} | {sum(a['dram__bytes_read.sum'])/max(1,len(t)):.2f} | {sum(a['dram__bytes_write.sum'])/max(1,len(t)):.2f} |")
