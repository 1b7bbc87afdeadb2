"""Summarise a SDPA_UMMA_TRACE timeline (clock64 stamps of CTA (0,0))."""
import collections
import statistics as st
import sys

T = collections.defaultdict(dict)
for l in open(sys.argv[1]).read().split('\n')[1:]:
    if l.strip():
        r, j, e, c = map(int, l.split())
        T[(r, j)][e] = c
t0 = min(v for d in T.values() for v in d.values())
names = {0: 'A-lo', 1: 'A-hi', 2: 'B-lo', 3: 'B-hi', 4: 'MMA', 5: 'TMA'}
for j in range(8, 11):
    for r in (0, 1, 2, 4):
        d = T.get((r, j), {})
        print(f"j={j:2d} {names[r]:5s}", " ".join(f"e{e}={d[e]-t0:7d}" for e in sorted(d)))


def avg(f):
    vals = [f(j) for j in range(4, 22)]
    vals = [v for v in vals if v is not None]
    return round(st.mean(vals)) if vals else None


def d(r, a, b, dj=0):
    def f(j):
        if (r, j) in T and (r, j + dj) in T and a in T[(r, j)] and b in T[(r, j + dj)]:
            return T[(r, j + dj)][b] - T[(r, j)][a]
        return None
    return f


for r in (0, 1, 2, 3):
    print(names[r], "s_full->1st chunk", avg(d(r, 0, 1)), "| chunks (ld+max+exp+st)", avg(d(r, 1, 2)), "| exchange", avg(d(r, 2, 3)),
          "| decide/redo", avg(d(r, 3, 4)), "| wait_st+arrive", avg(d(r, 4, 5)), "| total", avg(d(r, 0, 5)), "| period", avg(d(r, 0, 0, 1)),
          "| idle", avg(d(r, 5, 0, 1)))
print("MMA: wait pA", avg(d(4, 0, 1)), "| issue PV_A", avg(d(4, 1, 2)), "| issue S_A", avg(d(4, 2, 3)), "| wait pB", avg(d(4, 3, 4)),
      "| issue PV_B,S_B", avg(d(4, 4, 5)), "| period", avg(d(4, 0, 0, 1)))
