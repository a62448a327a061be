"""Attribute a kernel's executed warp instructions to code regions from an ncu report's SASS source page
(`ncu -i X.ncu-rep --page source --csv --print-source sass -k regex:KERNEL > src.csv; python scripts/ncu_regions.py src.csv`):
consecutive instructions with (almost) equal execution counts form a region -- loop bodies and per-warp scalar sections show up
with their share of the kernel's instructions and their FP64 fraction.  This is how the 212-instruction per-node prologue of
the setup kernel's n_eff loop (41 % of its instructions) was found."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[idx["Instructions Executed"]].strip().isdigit():
        continue  # short rows / repeated header blocks (one per matching kernel)
    ins.append((r[idx["Source"]].strip(), int(r[idx["Instructions Executed"]] or 0), int(r[idx["# Samples"]] or 0)))
tot = sum(e for _, e, _ in ins)
fp = re.compile(r"(@\S+\s+)?(DFMA|DMUL|DADD|DSETP|DMMA)")
regions, cur = [], None
for k, (s, e, n) in enumerate(ins):
    if cur and abs(e - cur["lvl"]) <= 0.02 * max(e, cur["lvl"]):
        cur["n"] += 1; cur["tot"] += e; cur["end"] = k; cur["samp"] += n
        cur["fp"] += e if fp.match(s) else 0
    else:
        if cur:
            regions.append(cur)
        cur = dict(lvl=e, n=1, tot=e, start=k, end=k, samp=n, fp=e if fp.match(s) else 0)
regions.append(cur)
regions.sort(key=lambda r: -r["tot"])
print("total executed warp instructions %d in %d SASS instructions" % (tot, len(ins)))
for r in regions[: int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print("instr %5d..%5d  n=%4d  exec/instr=%10d  share=%5.1f%%  FP64 in region=%3.0f%%  samples=%d"
          % (r["start"], r["end"], r["n"], r["lvl"], 100.0 * r["tot"] / tot, 100.0 * r["fp"] / max(1, r["tot"]), r["samp"]))
