"""Instruction mix + stall samples per opcode from an `ncu --page source --csv` dump (gzip)."""
import collections
import csv
import gzip
import re
import sys

rows = list(csv.reader(gzip.open(sys.argv[1], "rt")))
print(rows[0][1][:100])
hdr, data = rows[1], rows[2:]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp = collections.Counter(), collections.Counter()
for r in data:
    if len(r) <= max(iS, iE, iSamp) or not r[iE].strip().isdigit():
        continue
    src = r[iS].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src
    op = ".".join(op.split(".")[:2])
    ops[op] += int(r[iE])
    samp[op] += int(r[iSamp])
tot, ts = sum(ops.values()), sum(samp.values())
print("total warp-instr", tot, "samples", ts)
for op, c in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"  {op:22s} {c:12d} {100 * c / tot:5.1f}%  samples {100 * samp[op] / ts:5.1f}%")
