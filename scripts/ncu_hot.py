"""Top stall sites of one kernel from an .ncu-rep (source page, SASS view): python scripts/ncu_hot.py rep kernel-regex [n]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
si, ss = h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
ie = h.index("Instructions Executed")
body = [r for r in rows[hi + 1:] if len(r) > ss and r[0].startswith("0x")]
# only the first kernel instance
seen, first = set(), []
for r in body:
    if r[0] in seen: break
    seen.add(r[0]); first.append(r)
tot = sum(int(r[ss] or 0) for r in first)
print(f"{len(first)} SASS instructions, {tot} stall samples")
order = sorted(range(len(first)), key=lambda i: -int(first[i][ss] or 0))[:n]
for i in sorted(order):
    r = first[i]
    print(f"{i:5d} {int(r[ss] or 0):6d} {100*int(r[ss] or 0)/max(tot,1):5.1f}%  x{r[ie]:>8s}  {r[si].strip()[:110]}")
