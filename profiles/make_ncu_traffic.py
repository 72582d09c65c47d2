"""profiles/ncu_traffic.json from the committed `ncu --set full` summaries (profiles/r*_ncu_full_*.csv, written by
summarize_ncu_full.py): DRAM bytes (read + write) per launch of every kernel captured, newest round wins.
bench.py reads roofline.traffic from this file instead of a literal.   python profiles/make_ncu_traffic.py"""
import csv, glob, json, os, re
here = os.path.dirname(os.path.abspath(__file__))
unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for path in sorted(glob.glob(os.path.join(here, "r*_ncu_full_*.csv"))):
    rows = list(csv.reader(open(path)))
    if not rows:
        continue
    hdr = rows[0]
    col = {}
    for i, h in enumerate(hdr):
        m = re.match(r"(\w+)(?:\[(.*)\])?$", h)
        if m:
            col[m.group(1)] = (i, m.group(2) or "")
    if "dram_read" not in col or "kernel" not in col:
        continue
    acc = {}
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"^void\s+", "", r[col["kernel"][0]]).split("<")[0].strip()
        try:
            rd = float(r[col["dram_read"][0]]) * unit.get(col["dram_read"][1], 1.0)
            wr = float(r[col["dram_write"][0]]) * unit.get(col["dram_write"][1], 1.0) if "dram_write" in col else 0.0
            dur = float(r[col["dur_us"][0]]) if "dur_us" in col else 0.0
        except ValueError:
            continue
        a = acc.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1; a[1] += rd + wr; a[2] += dur
    for name, (n, b, d) in acc.items():
        out[name] = {"dram_bytes_per_launch": b / n, "launches_captured": n, "us_per_launch_under_ncu": d / n,
                     "source": "profiles/" + os.path.basename(path)}
json.dump(out, open(os.path.join(here, "ncu_traffic.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
