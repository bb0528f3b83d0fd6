"""Summarise an `ncu --page source --csv` dump of favor_tc_kernel: stall samples per mbarrier wait."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = rows[2:]
tot = sum(int(r[isamp]) for r in data if r[isamp].isdigit())
print("total samples", tot, "warp-instructions", sum(int(r[iex]) for r in data if r[iex].isdigit()))
names = {0x00: 'tfull0', 0x08: 'tfull1', 0x10: 'tfull2', 0x18: 'tempty0', 0x20: 'tempty1', 0x28: 'tempty2', 0x30: 'ufull0',
         0x38: 'ufull1', 0x40: 'ufree0', 0x48: 'ufree1', 0x50: 'fready0', 0x58: 'fready1', 0x60: 'ffree0', 0x68: 'ffree1',
         0x70: 'd3full0', 0x78: 'd3full1', 0x80: 'd3free0', 0x88: 'd3free1', 0x90: 'ctxfull', 0x98: 'ctxready'}
agg = {}
for i, r in enumerate(data):
    s = r[isrc]
    if 'TRYWAIT' in s and r[iex].isdigit() and int(r[iex]) > 0:
        m = re.search(r'\+0x35(0[0-9a-f]{2})\]', s)
        off = int(m.group(1), 16) if m else -1
        nxt = sum(int(data[j][isamp]) for j in range(i, min(i + 4, len(data))))
        print(f"#{i:4d} ex={r[iex]:>9s} samples~{nxt:6d} {names.get(off, '?'):8s} {s[:70]}")
# busiest non-wait instructions
lst = sorted(((int(r[isamp]), i) for i, r in enumerate(data) if r[isamp].isdigit()), reverse=True)[:25]
for smp, i in lst:
    print(f"   #{i:4d} {smp:6d} ex={data[i][iex]:>9s} {data[i][isrc][:90]}")
