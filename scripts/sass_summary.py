"""Per-kernel tensor-core / TMA instruction evidence from the built library: counts of the SASS mnemonics that prove
tcgen05 (UTCHMMA / UTCQMMA ...), TMEM loads (LDTM), TMA (UTMALDG), legacy mma.sync (HMMA) in every kernel of libdbmm.so.
    python scripts/sass_summary.py > profiles/r2_sass_summary.md"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "debiasing-multi-modal_b200", "libdbmm.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, counts = None, collections.OrderedDict()
pat = {"UTCHMMA (tcgen05.mma f16/tf32)": r"\bUTC[HQ]MMA", "LDTM (tcgen05.ld)": r"\bLDTM", "UTMALDG (TMA load)": r"\bUTMALDG",
       "UTCBAR (tcgen05.commit)": r"\bUTCBAR", "HMMA (mma.sync)": r"\bHMMA", "SYNCS (mbarrier)": r"\bSYNCS", "RED/ATOM": r"\b(RED|ATOM)[G.]"}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("dbmm::", "").replace("void ", "")
        counts[kern] = collections.Counter()
        continue
    if kern:
        for k, p in pat.items():
            if re.search(p, line):
                counts[kern][k] += 1
print("# SASS instruction summary of libdbmm.so (sm_100a)\n")
print("`cuobjdump -sass debiasing-multi-modal_b200/libdbmm.so`, counted per kernel by `scripts/sass_summary.py`.\n")
cols = list(pat)
print("| kernel | " + " | ".join(cols) + " |\n|---|" + "---|" * len(cols))
for k, c in counts.items():
    if any(c[x] for x in cols[:5]):
        print(f"| `{k[:70]}` | " + " | ".join(str(c[x]) for x in cols) + " |")
