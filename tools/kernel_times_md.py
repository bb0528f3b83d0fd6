"""Markdown table (time, achieved rate, fraction of the measured peak) from the output of tools/bench_gemm_shapes.py,
bench_favor.py and bench_pair_misc.py. usage: python tools/kernel_times_md.py log [MEASURED_PEAKS.json]"""
import json, re, sys
log = open(sys.argv[1]).read().splitlines()
peaks = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else {}
tf, hbm = peaks.get("bf16_tflops_sustained", 1377.7), peaks.get("hbm_gbs", 6533.0)
print("| kernel shape | time (us) | achieved | of peak |\n|---|---:|---:|---|")
for ln in log:
    m = re.match(r"(.+?)\s+([\d.]+) us\s+([\d.]+) (TFLOP/s|GB/s)", ln)
    if not m:
        continue
    name, us, rate, unit = m.group(1).strip(), float(m.group(2)), float(m.group(3)), m.group(4)
    pk = tf if unit == "TFLOP/s" else hbm
    print(f"| {name} | {us:.1f} | {rate:.0f} {unit} | {100 * rate / pk:.0f} % of {pk:.0f} {unit} |")
