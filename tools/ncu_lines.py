"""Joins an ncu source-page CSV (SASS, per-instruction samples) with nvdisasm -g line info: stall samples and executed
instructions per source line / per opcode.  usage: ncu_lines.py <src.csv> <kernel.sass (nvdisasm -g)> [top]"""
import collections
import csv
import re
import sys

src_csv, sass = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
line_of, cur = {}, None
for l in open(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
base = None
by_line, by_line_i, by_op, by_op_i = (collections.Counter() for _ in range(4))
tot = toti = 0
for r in rows[2:]:
    try:
        a, n, s = int(r[ia], 16), int(r[iex]), int(r[ismp])
    except ValueError:
        continue
    base = a if base is None else base
    k = line_of.get(a - base)
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
    op = m.group(2) if m else "?"
    op = op if op.split(".")[0] in ("LDS", "STS", "LDG", "SHFL", "MUFU", "LDL", "STL", "BAR", "SYNCS") else op.split(".")[0]
    by_line[k] += s; by_line_i[k] += n; by_op[op] += s; by_op_i[op] += n
    tot += s; toti += n
print(f"total warp instructions {toti}, samples {tot}")
print("-- by opcode (instr %, samples %)")
for k, v in by_op_i.most_common(28):
    print(f"  {k:26s} {v / toti * 100:6.2f}  {by_op[k] / tot * 100:6.2f}")
print("-- by source line (samples %, instr %)")
for k, v in by_line.most_common(top):
    print(f"  {str(k):44s} {v / tot * 100:6.2f}  {by_line_i[k] / toti * 100:6.2f}")
