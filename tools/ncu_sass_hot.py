"""Per-SASS-instruction view of an ncu report (one kernel): executed-weighted opcode histogram and the hottest
contiguous address ranges.   python tools/ncu_sass_hot.py report.ncu-rep [--ranges N] [--dump out.txt]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
iS, iE, iSt = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
ins = [(r[iS].strip(), int(r[iE] or 0), int(r[iSt] or 0)) for r in rows[hdr + 1:] if len(r) > iE]
tot = sum(e for _, e, _ in ins); tots = sum(s for _, _, s in ins)
print("static instructions", len(ins), "executed", tot, "stall samples", tots)
hist = collections.Counter(); sh = collections.Counter()
for s, e, st in ins:
    t = s.split()
    op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
    hist[op.split(".")[0]] += e; sh[op.split(".")[0]] += st
print("executed by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in hist.most_common(24)))
print("stall samples by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(tots, 1)) for k, v in sh.most_common(16)))
if "--dump" in sys.argv:
    with open(sys.argv[sys.argv.index("--dump") + 1], "w") as f:
        for i, (s, e, st) in enumerate(ins):
            f.write("%5d %12d %7d  %s\n" % (i, e, st, s))
# hottest ranges: split where the executed count changes by more than 2x
n = int(sys.argv[sys.argv.index("--ranges") + 1]) if "--ranges" in sys.argv else 12
ranges = []; a = 0
for i in range(1, len(ins) + 1):
    if i == len(ins) or not (0.5 <= (ins[i][1] + 1) / (ins[a][1] + 1) <= 2.0):
        ranges.append((sum(e for _, e, _ in ins[a:i]), a, i)); a = i
for w, a, b in sorted(ranges, reverse=True)[:n]:
    print("  [%5d,%5d) %5.1f%% of executed, %d static, ~%d exec each" % (a, b, 100.0 * w / tot, b - a, w // max(b - a, 1)))
