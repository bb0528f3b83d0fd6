"""Print a window of the FAVOR kernel's developer timeline (RFK_FAVOR_TRACE=file): clock64 deltas of
CTA 0 for the U issuer, the consumer issuer and two feature warps."""
import collections, sys
ev = collections.defaultdict(list)
for l in open(sys.argv[1]):
    r, e, c = l.split()
    ev[int(r)].append((int(e), int(c)))
t0 = min(v[0][1] for v in ev.values())
allv = sorted((c - t0, r, e) for r, v in ev.items() for e, c in v)
names = {10: 'U:ufree ok c0', 11: 'U:ufree ok c1', 12: 'U:ufree ok c2', 13: 'U:tile ok', 14: 'U:issued', 20: 'Ck:start c0',
         21: 'Ck:start c1', 22: 'Ck:start c2', 23: 'Ck:fready ok', 24: 'Ck:issued', 30: 'Cq:start c0', 31: 'Cq:start c1',
         32: 'Cq:start c2', 33: 'Cq:fready ok', 34: 'Cq:issued', 40: 'F:start c0', 41: 'F:start c1', 42: 'F:start c2',
         43: 'F:ufull ok', 44: 'F:loaded+ufree', 45: 'F:math done', 46: 'F:ffree ok', 47: 'F:stored', 48: 'F:fenced',
         49: 'F:fready arrived', 50: 'E:epi start', 51: 'E:d3full ok', 60: 'R:readout start', 61: 'R:ctxfull ok',
         62: 'R:ctxready arrived'}
lo = allv[len(allv) // 2][0] if len(sys.argv) < 3 else int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 120
cnt = 0
for c, r, e in allv:
    if c >= lo and cnt < n:
        print(f"{c - lo:7d}  role{r} {'  ' * r * 4}{names.get(e, e)}")
        cnt += 1
