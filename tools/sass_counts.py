"""Per-kernel counts of the Blackwell-native SASS mnemonics in librfk.so (run on the CPU box):
python tools/sass_counts.py > profiles/r02_sass.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "rosettafold-pytorch_b200", "librfk.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCHMMA tmem[", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "HMMA"]
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::|rfk::|<unnamed>::", "", cur)
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in KEYS:
        if k == "UTCHMMA tmem[":
            if re.search(r"UTCHMMA tmem\[", line):
                counts[cur][k] += 1
        elif re.search(r"\b" + k, line):
            counts[cur][k] += 1
print("# Round 2 — SASS evidence (`cuobjdump -sass rosettafold-pytorch_b200/librfk.so`, sm_100a)\n")
print("tcgen05.mma -> `UTCHMMA` (`UTCHMMA tmem[..]` = A operand read from tensor memory), tcgen05.ld / tcgen05.st -> `LDTM` / `STTM`,")
print("TMA loads / stores / L2 prefetch -> `UTMALDG` / `UTMASTG` / `UTMAPF`, tcgen05.commit -> `UTCBAR`, mbarrier -> `SYNCS`; `HMMA` would be")
print("the legacy mma.sync path (absent). Regenerate with `python tools/sass_counts.py`.\n")
print("| kernel | " + " | ".join(f"`{k}`" for k in KEYS) + " |")
print("|---|" + "---:|" * len(KEYS))
tot = collections.Counter()
for name, c in counts.items():
    if not any(c[k] for k in KEYS if k != "SYNCS"):
        continue
    print(f"| `{name[:70]}` | " + " | ".join(str(c[k]) for k in KEYS) + " |")
    tot.update(c)
print("| **total** | " + " | ".join(f"**{tot[k]}**" for k in KEYS) + " |")
print(f"\n{len(counts)} kernels in the library; kernels without any of these instructions (SIMT elementwise / fp32 validation kernels) are not listed.")
