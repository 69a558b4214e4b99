"""Summarise an `ncu --page source --csv --print-source sass` dump: executed-instruction
histogram by opcode, stall-sample histogram, and the hottest SASS lines."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address" and r[1] == "Source"]
h = rows[hi[0]]
end = hi[1] - 1 if len(hi) > 1 else len(rows)
body = [r for r in rows[hi[0] + 1:end] if len(r) == len(h)]
ix = {n: i for i, n in enumerate(h)}
ops, total = collections.Counter(), 0
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
stalls = collections.Counter()
samples_by_line = []
for r in body:
    n = int(r[ix["Instructions Executed"]] or 0)
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    op = m.group(2) if m else "?"
    ops[op] += n
    total += n
    for c in stall_cols:
        stalls[c] += int(r[ix[c]] or 0)
    samples_by_line.append((int(r[ix["# Samples"]] or 0), n, r[ix["Source"]][:90]))
chunks = max(int(b[ix["Instructions Executed"]] or 0) for b in body)
print("total warp-instructions executed:", total, " max per-line count:", chunks, " => instr per hottest-loop pass:", round(total / max(chunks, 1), 1))
print("by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / total) for k, v in ops.most_common(28)))
ts = sum(stalls.values())
print("stall samples:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / ts) for k, v in stalls.most_common(10)))
print("hottest lines by stall samples:")
for s, n, src in sorted(samples_by_line, reverse=True)[:top_n]:
    print("  %6d samples  %9d exec  %s" % (s, n, src))
