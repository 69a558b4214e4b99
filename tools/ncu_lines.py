"""Warp-instructions and stall samples per SOURCE line of a kernel, from `ncu -i rep --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_lines.py report.ncu-rep [chunks-per-launch] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
chunks = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = hdr = None
agg, samp = {}, {}
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) == 2:
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) // 2:
        continue
    try:
        ln, n, s = int(r[0]), int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    if n:
        k = (cur, ln, r[1].strip()[:100])
        agg[k] = agg.get(k, 0) + n
        samp[k] = samp.get(k, 0) + s
tot = sum(agg.values())
print("warp-instructions %d  = %.1f per chunk; stall samples %d" % (tot, tot / chunks, sum(samp.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print("%7.1f %6d  %s:%d  %s" % (v / chunks, samp[k], k[0], k[1], k[2]))
