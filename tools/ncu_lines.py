"""Stall samples of an ncu report aggregated per SOURCE LINE of the kernel's .cu file (the outermost frame of every
inlined helper), read on the CPU box: `ncu --page source` gives samples per SASS instruction, `nvdisasm
--print-line-info-inline` of the same object gives the line of every instruction; both list the function in order.
usage: python tools/ncu_lines.py rep.ncu-rep build/obj.o kernel_name_substring [min_percent]"""
import csv, os, re, subprocess, sys, tempfile

rep, obj, kname = sys.argv[1:4]
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
sass = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[idx["# Samples"]])
    except ValueError:
        continue
    sass.append((s, r[idx["Source"]], {h: int(r[idx[h]] or 0) for h in stall_cols}))

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# RFK_LINES_INNER=1: attribute to the INNERMOST frame that lies in a .cu file (lambda bodies) instead of the outermost
inner = os.environ.get("RFK_LINES_INNER") is not None
lines, cur, infn, block = [], None, False, []
for ln in dis.splitlines():
    if ln.startswith(".text."):
        infn = kname in ln
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        block.append((m.group(1), int(m.group(2)), "inlined at" in m.group(3)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
        if block:
            if inner:
                cu = [b for b in block if b[0].endswith(".cu")]
                cur = (cu[0][0], cu[0][1]) if cu else (block[-1][0], block[-1][1])
            else:
                outer = [b for b in block if not b[2]]
                cur = (outer[-1][0], outer[-1][1]) if outer else (block[-1][0], block[-1][1])
            block = []
        lines.append(cur)
if len(lines) != len(sass):
    print(f"warning: {len(lines)} instructions in the object vs {len(sass)} in the report")
total = sum(s for s, _, _ in sass)
agg = {}
for (s, _, st), loc in zip(sass, lines):
    a = agg.setdefault(loc, [0, {}])
    a[0] += s
    for k, v in st.items():
        a[1][k] = a[1].get(k, 0) + v
srcs = {}
print(f"total samples {total}")
for loc in sorted(agg, key=lambda l: (l is None, l)):
    s, st = agg[loc]
    if s < total * minpct / 100:
        continue
    f, n = loc if loc else ("?", 0)
    if f not in srcs:
        try:
            srcs[f] = open(f).read().splitlines()
        except OSError:
            srcs[f] = []
    text = srcs[f][n - 1].strip() if 0 < n <= len(srcs[f]) else ""
    top = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
    print(f"{100*s/total:5.1f}% {os.path.basename(f)}:{n:<4d} {text[:80]:80s} [{top}]")
