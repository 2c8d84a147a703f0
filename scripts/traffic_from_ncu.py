"""profiles/r1_traffic.json from an ncu launch list (metrics dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum,
default cache control = cold caches): mean DRAM bytes per launch of every kernel.  usage: traffic_from_ncu.py in.csv [more.csv] out.json"""
import collections, csv, json, sys
*ins, out = sys.argv[1:]
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for path in ins:
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Metric Unit")
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}
    for r in rows[h + 1:]:
        if len(r) > vi:
            name = r[ki].split("(")[0].replace("void ", "").split("<")[0].replace("dbmm::", "")
            agg[name][r[mi]].append(float(r[vi].replace(",", "")) * mult.get(r[ui], 1))
res = {}
for k, m in agg.items():
    rd, wr = m.get("dram__bytes_read.sum", [0]), m.get("dram__bytes_write.sum", [0])
    res[k] = round((sum(rd) / len(rd)) + (sum(wr) / len(wr)))
    res[k + ":detail"] = {"launches": len(rd), "dram_read_bytes": sum(rd) / len(rd), "dram_write_bytes": sum(wr) / len(wr),
                          "ncu_duration_us": sum(m.get("gpu__time_duration.sum", [0])) / max(len(m.get("gpu__time_duration.sum", [0])), 1)}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if not k.endswith(":detail")}, indent=1))
