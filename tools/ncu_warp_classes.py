"""Split the warp-state samples of an `ncu --page source --csv` dump by how often each SASS instruction ran:
in the pipelined sweeps the warps of a tile take different code (table chunks / constant chunks), so the
execution count of an instruction says which warps run it.  python tools/ncu_warp_classes.py source.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
heads = [i for i, r in enumerate(rows) if "# Samples" in r]
for n, hi in enumerate(heads):
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    end = heads[n + 1] - 1 if n + 1 < len(heads) else len(rows)
    seg = [r for r in rows[hi + 1:end] if len(r) >= len(hdr) and r[ix["# Samples"]].strip().isdigit()]
    cat, cnt, lsb, bar = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    for r in seg:
        ex = int(r[ix["Instructions Executed"]] or 0)
        cat[ex] += int(r[ix["# Samples"]]); cnt[ex] += 1
        lsb[ex] += int(r[ix["stall_long_sb"]] or 0); bar[ex] += int(r[ix["stall_barrier"]] or 0)
    tot = sum(cat.values())
    print("kernel %d: %d samples" % (n, tot))
    for ex, s in sorted(cat.items(), key=lambda kv: -kv[1])[:8]:
        print("  executed %8d x  %5d instructions  %6d samples (%4.1f%%)  long scoreboard %5d  barrier %5d" % (ex, cnt[ex], s, 100.0 * s / tot, lsb[ex], bar[ex]))
    top = max(cat)  # instructions every warp runs
    for ex in sorted(cat, reverse=True):
        if ex in (top,) or cat[ex] < 0.01 * tot or ex > top:
            continue
        regions, cur = [], None
        for i, r in enumerate(seg):
            if int(r[ix["Instructions Executed"]] or 0) == ex:
                if cur is None or i - cur[1] > 8:
                    cur = [i, i]; regions.append(cur)
                cur[1] = i
        for a, b in regions:
            sub = [r for r in seg[a:b + 1] if int(r[ix["Instructions Executed"]] or 0) == ex]
            smp = sum(int(r[ix["# Samples"]]) for r in sub)
            if smp < 0.002 * tot:
                continue
            l = sum(int(r[ix["stall_long_sb"]] or 0) for r in sub)
            ops = collections.Counter((r[ix["Source"]].split()[1] if r[ix["Source"]].startswith("@") else r[ix["Source"]].split()[0]).split(".")[0] for r in sub)
            print("    x%-7d SASS %4d-%4d  %4d instr  %5d samples  long sb %5d  %s" % (ex, a, b, len(sub), smp, l, dict((k, v) for k, v in ops.items() if v > 10)))
