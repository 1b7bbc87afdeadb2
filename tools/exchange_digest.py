"""Digest of an SDPA_EXCHANGE_TRACE run (one file per rank, %globaltimer stamps in ns, see csrc/merge_kernels.cu):
per exchange step, how long the root waited for the slowest shard (skew), how long the flag hop took, how long the merge
kernel ran.  Usage: python tools/exchange_digest.py gpurun_out/xtrace [nranks]"""
import glob
import statistics as st
import sys

base = sys.argv[1]
files = sorted(glob.glob(base + ".rank*"), key=lambda p: int(p.rsplit("rank", 1)[1]))
ranks = {}
for p in files:
    r = int(p.rsplit("rank", 1)[1])
    rows = {}
    for line in open(p).read().splitlines()[1:]:
        e, pub, mb, fs, md = map(int, line.split())
        rows[e] = (pub, mb, fs, md)
    ranks[r] = rows
if 0 not in ranks:
    sys.exit("no trace of rank 0 under " + base)
common = set(ranks[0])
for r in ranks:
    common &= set(ranks[r])
steps = sorted(e for e in common if all(ranks[r][e][0] for r in ranks) and ranks[0][e][3])
steps = steps[len(steps) // 4:]          # drop the warm-up quarter
if not steps:
    sys.exit("no complete steps")


def med(xs):
    xs = list(xs)
    return st.median(xs) / 1e3 if xs else float("nan")


pub = {e: [ranks[r][e][0] for r in sorted(ranks)] for e in steps}
print(f"{len(ranks)} ranks, {len(steps)} exchange steps (us, medians; %globaltimer of different GPUs agrees to a few us)")
print(f"  publish skew (last - first shard's state published)        : {med(max(pub[e]) - min(pub[e]) for e in steps):8.1f}")
print(f"  root published -> root merge kernel started                 : {med(ranks[0][e][1] - ranks[0][e][0] for e in steps):8.1f}")
print(f"  last publish -> root saw every flag                         : {med(ranks[0][e][2] - max(pub[e]) for e in steps):8.1f}")
print(f"  root merge kernel: flags seen -> merged + consumed raised   : {med(ranks[0][e][3] - ranks[0][e][2] for e in steps):8.1f}")
print(f"  root merge kernel total (start -> done)                     : {med(ranks[0][e][3] - ranks[0][e][1] for e in steps):8.1f}")
per = [med(ranks[0][b][3] - ranks[0][a][3] for a, b in zip(steps, steps[1:]) if b == a + 1)]
print(f"  step period (done -> done of consecutive epochs)            : {per[0]:8.1f}")
late = [max(range(len(pub[e])), key=lambda i: pub[e][i]) for e in steps]
print("  slowest shard histogram (rank: steps)                        :", {r: late.count(r) for r in sorted(set(late))})
