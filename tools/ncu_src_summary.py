"""Summarise an `ncu --page source --csv` dump: stall reasons, opcode mix, hottest SASS lines."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); ops = collections.Counter(); opsamp = collections.Counter()
lines = []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ix["# Samples"]].strip().isdigit(): continue
    if len(r) < len(hdr): continue
    src = r[ix["Source"]]; ex = int(r[ix["Instructions Executed"]] or 0); smp = int(r[ix["# Samples"]] or 0)
    op = src.split()[0] if src and not src.startswith("@") else (src.split()[1] if len(src.split()) > 1 else src)
    op = op.split(".")[0]
    ops[op] += ex; opsamp[op] += smp
    for c in stall_cols: tot[c] += int(r[ix[c]] or 0)
    lines.append((smp, ex, src))
S = sum(tot.values())
print("stall samples:", S)
for k, v in tot.most_common(10): print("  %-28s %6.1f%%" % (k, 100.0 * v / S))
E = sum(ops.values())
print("warp instructions executed:", E)
for k, v in ops.most_common(22): print("  %-10s %6.1f%%  samples %5.1f%%" % (k, 100.0 * v / E, 100.0 * opsamp[k] / max(1, sum(opsamp.values()))))
print("hottest lines:")
for smp, ex, src in sorted(lines, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("  %6d smp %10d ex  %s" % (smp, ex, src[:110]))
