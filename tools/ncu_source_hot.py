"""Per-source-line stall samples of one kernel of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_source_hot.py REPORT.ncu-rep LAUNCH_INDEX [top]"""
import csv, subprocess, sys, io, collections

rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hdr]
ci = {n: i for i, n in enumerate(h)}
samp = ci["# Samples"]
stall_cols = [(n, i) for n, i in ci.items() if n.startswith("stall_") and "Not Issued" not in n]
print(rows[1][1][:100] if len(rows) > 1 else "")
lines = collections.OrderedDict()
cur = None
file_of = None
for r in rows[hdr + 1:]:
    if len(r) < len(h):
        if r and r[0] == "File Path": file_of = r[1].split("/")[-1]
        continue
    if r[0] != "":
        cur = (file_of, r[0], r[1].strip()[:90])
        lines.setdefault(cur, [0, collections.Counter(), 0])
        continue  # the per-line row aggregates its SASS rows below; count SASS rows only
    if cur is None: continue
    try: n = int(r[samp])
    except ValueError: continue
    e = lines[cur]
    e[0] += n
    e[2] += int(r[ci["Instructions Executed"]] or 0)
    for name, i in stall_cols:
        try: e[1][name] += int(r[i])
        except ValueError: pass
tot = sum(e[0] for e in lines.values()) or 1
print(f"total samples {tot}")
agg = collections.Counter()
for e in lines.values(): agg.update(e[1])
print("stalls:", ", ".join(f"{k[6:]} {100*v/tot:.0f}%" for k, v in agg.most_common(8)))
for k, e in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{n[6:]} {v}" for n, v in e[1].most_common(3))
    print(f"{100*e[0]/tot:5.1f}%  {k[0]}:{k[1]:>4}  inst {e[2]:>8}  [{st}]  {k[2]}")
