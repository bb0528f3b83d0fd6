"""Top SASS instructions by stall samples from `ncu -i rep --page source --csv` (read on the CPU box).
usage: python tools/ncu_hot.py rep.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[idx["# Samples"]])
    except ValueError:
        continue
    st = {h: int(r[idx[h]] or 0) for h in stall_cols}
    data.append((s, r[idx["Address"]], r[idx["Source"]], int(r[idx["Instructions Executed"]] or 0), st))
total = sum(d[0] for d in data)
print(f"total samples {total}")
agg = {}
for d in data:
    for k, v in d[4].items():
        agg[k] = agg.get(k, 0) + v
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > total * 0.01})
for i, d in enumerate(data):
    data[i] = d + (i,)
for s, addr, src, ex, st, i in sorted(data, key=lambda d: -d[0])[:n]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{s:7d} {100*s/total:5.1f}%  #{i:4d} x{ex:8d}  {src[:90]:90s} {top}")
