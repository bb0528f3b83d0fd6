"""Per-kernel shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list (read on the CPU box).
usage: python tools/launch_summary.py launches.csv [prefix] > summary.md   (prefix, e.g. rfk:: : kernels whose name does not
start with it are summed on one line instead of being listed)"""
import csv, re, sys
path = sys.argv[1]
prefix = sys.argv[2] if len(sys.argv) > 2 else ""
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg, total, n = {}, 0.0, 0
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(anonymous namespace\)|<unnamed>", "", r[ix["Kernel Name"]]).replace("::::", "::")
    name = name.split("(")[0].replace("void ", "")
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
    if prefix and not name.startswith(prefix):
        name = f"(not {prefix}*)"
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
    total += us
    n += 1
print(f"{n} launches, {total / 1e3:.2f} ms summed\n")
print("| kernel | launches | sum (us) | avg (us) | share |\n|---|---:|---:|---:|---:|")
for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name[:80]}` | {c} | {us:.1f} | {us / c:.1f} | {100 * us / total:.1f}% |")
fam = {}
for name, (c, us) in agg.items():
    f = ("FAVOR" if "favor" in name else "GEMM (tcgen05)" if "gemm_tc" in name else "GEMM (SIMT fp32)" if "gemm_f32" in name
         else "LayerNorm" if "layernorm" in name else "not librfk" if name.startswith("(not") else "other librfk kernels")
    fam[f] = fam.get(f, 0.0) + us
print("\n| family | sum (us) | share |\n|---|---:|---:|")
for f, us in sorted(fam.items(), key=lambda kv: -kv[1]):
    print(f"| {f} | {us:.1f} | {100 * us / total:.1f}% |")
