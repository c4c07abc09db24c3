"""Splits an ncu --page source --csv SASS listing at BAR.SYNC instructions and prints, per segment, the share of
warp-stall samples, the warp instructions executed, and the first FFMA/LDG-ish hints.  usage: ncu_phases.py src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iSamp, iExec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
segs, cur = [], {"n": 0, "samp": 0, "exec": 0, "ffma": 0, "first": None, "ops": {}}
for r in rows[2:]:
    if len(r) <= iExec:
        continue
    src = r[iS].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    cur["n"] += 1
    cur["samp"] += int(r[iSamp] or 0)
    e = int(r[iExec] or 0)
    cur["exec"] += e
    cur["ops"][op] = cur["ops"].get(op, 0) + e
    if "BAR.SYNC" in src:
        segs.append(cur)
        cur = {"n": 0, "samp": 0, "exec": 0, "ffma": 0, "first": None, "ops": {}}
segs.append(cur)
tot = sum(s["samp"] for s in segs)
tote = sum(s["exec"] for s in segs)
print(f"total samples {tot}, warp instructions {tote}")
for i, s in enumerate(segs):
    top = sorted(s["ops"].items(), key=lambda kv: -kv[1])[:4]
    print(f"seg {i:2d}: static {s['n']:5d}  samples {100*s['samp']/tot:5.1f}%  exec {100*s['exec']/tote:5.1f}%  "
          f"samples/exec {s['samp']/max(s['exec'],1)*1e3:6.2f}  " + " ".join(f"{k}:{100*v/max(s['exec'],1):.0f}%" for k, v in top))
