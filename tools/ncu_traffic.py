"""profiles/traffic.json from an ncu --set full report of tools/prof_r2.py: dram__bytes_read.sum + dram__bytes_write.sum per launch
of the kernels bench.py reports a roofline for.  usage: python tools/ncu_traffic.py report.ncu-rep [capture label]"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(rep)
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


out = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200::", "").replace("ndt::", "").strip()
    base = name.split("<")[0]
    b = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    dur = float(r[ix["gpu__time_duration.sum"]].replace(",", ""))
    dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(units[ix["gpu__time_duration.sum"]], 1.0)
    out.setdefault(base, []).append({"dram_bytes": b, "duration_us": dur_us, "grid": r[ix["launch__grid_size"]], "id": r[ix["ID"]]})
commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
kern = {}
alias = {"k_search_w": "k_search", "k_ndt_score_batch": "k_ndt_score_batch", "k_ndt_eval": "k_ndt_eval", "k_obs": "k_obs", "k_knn5_w": "k_knn5"}
for base, launches in out.items():
    if base not in alias:
        continue
    big = max(launches, key=lambda l: l["duration_us"])   # the working launch (search kernels of passes that reuse neighbours are no-ops)
    kern[alias[base]] = {"dram_bytes_per_launch": int(big["dram_bytes"]), "duration_us_under_ncu": big["duration_us"], "grid": big["grid"],
                         "launches_in_capture": len(launches), "capture": label}
json.dump({"commit": commit, "how": "ncu --set full --clock-control none on tools/prof_r2.py (L2 flushed before every measured call)", "kernels": kern},
          open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(kern, indent=1))
