"""Dev: per-opcode and per-source-line instruction / stall breakdown of one kernel from `ncu --page source --csv` output.
usage: ncu -i rep --page source --csv --kernel-name regex:K [--launch-skip n --launch-count 1] > src.csv; python tools/ncu_src.py src.csv queries"""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
nq = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
seen, data = set(), []
for r in rows[2:]:
    if len(r) > 5 and r[0].startswith("0x") and r[0] not in seen:
        seen.add(r[0]); data.append(r)
def g(r, k):
    try: return int(float(r[ix[k]] or 0))
    except Exception: return 0
tot = sum(g(r, "# Samples") for r in data)
agg = {}
for r in data:
    t = r[ix["Source"]].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    a = agg.setdefault(op, [0, 0, 0, 0, 0, 0])
    a[0] += g(r, "# Samples"); a[1] += g(r, "Instructions Executed"); a[2] += g(r, "stall_mio"); a[3] += g(r, "stall_short_sb")
    a[4] += g(r, "stall_long_sb"); a[5] += g(r, "Thread Instructions Executed")
print("instr/query %.1f  samples %d" % (sum(a[1] for a in agg.values()) / nq, tot))
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{op:10s} exec/q {a[1]/nq:7.1f} lanes {a[5]/max(a[1],1):5.1f} samples {100*a[0]/max(tot,1):5.1f}% mio {a[2]:6d} short_sb {a[3]:6d} long_sb {a[4]:6d}")
# cumulative executed count along the address order, to find hot regions
acc, marks = 0, []
for i, r in enumerate(data):
    acc += g(r, "Instructions Executed")
    if i % 64 == 63 or i == len(data) - 1:
        marks.append((i, acc / nq))
print("cumulative instr/query by static index:", [(i, round(v, 1)) for i, v in marks])
