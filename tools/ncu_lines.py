"""Per-source-line hot spots of an ncu report: python tools/ncu_lines.py report.ncu-rep [launch_index] [top]
(needs -lineinfo and --import-source on; uses the cuda,sass correlated source page)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
launches, cur, fpath, hdr = [], None, None, None
seen_files = set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1]
        continue
    if r[0] == "Function Name":
        key = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        if cur is None or (fpath, key) in seen_files:
            cur = []
            launches.append(cur)
            seen_files = set()
        seen_files.add((fpath, key))
        continue
    if r[0].isdigit() and hdr:
        cur.append((fpath.split("/")[-1], int(r[0]), r[1], dict(zip(hdr[4:], r[4:]))))
L = launches[which]
def f(d, k):
    try:
        return float(d.get(k, 0) or 0)
    except ValueError:
        return 0.0
tot_i = sum(f(d, "Instructions Executed") for _, _, _, d in L)
tot_s = sum(f(d, "# Samples") for _, _, _, d in L)
print(f"launches {len(launches)}; launch {which}: warp instructions {tot_i:.0f}, samples {tot_s:.0f}")
print("%-14s %5s %7s %7s %6s %6s %6s %6s  %s" % ("file", "line", "inst%", "samp%", "thr", "longsb", "wait", "mio", "source"))
for fn, ln, src, d in sorted(L, key=lambda x: -f(x[3], "# Samples"))[:top]:
    print("%-14s %5d %7.2f %7.2f %6.1f %6.0f %6.0f %6.0f  %s" % (fn, ln, 100 * f(d, "Instructions Executed") / tot_i, 100 * f(d, "# Samples") / tot_s,
          f(d, "Avg. Threads Executed"), f(d, "stall_long_sb"), f(d, "stall_wait"), f(d, "stall_mio"), src.strip()[:110]))
