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
        v = list(map(int, line.split()))
        rows[v[0]] = v[1:]
    ranks[r] = rows
if 0 not in ranks:
    sys.exit("no trace of rank 0 under " + base)
R = ranks[0]
n = len(ranks)
# in-stream root merge (SDPA_ROOT_MERGE=instream): the root publishes nothing (slot 0 stays 0) and waits for n-1 flags
instream = sum(1 for e in R if R[e][3] and not R[e][0]) > len(R) // 2
nflags = min(n - 1 if instream else n, 8)
steps = sorted(e for e in R if R[e][3] and all(R[e][4 + r] for r in range(nflags)))
steps = steps[len(steps) // 4:]          # drop the warm-up quarter
if not steps:
    sys.exit("no complete steps")


def med(xs):
    xs = list(xs)
    return st.median(xs) / 1e3 if xs else float("nan")


# Every GPU has its own %globaltimer: all cross-rank times are taken on the ROOT's clock (when its merge kernel saw a flag).
seen = {e: [R[e][4 + r] for r in range(nflags)] for e in steps}
print(f"{n} ranks, {len(steps)} exchange steps (us, medians, root GPU's clock)" + (", in-stream root merge" if instream else ""))
if not instream:
    print(f"  root's own state published -> root merge kernel started            : {med(R[e][1] - R[e][0] for e in steps):8.1f}")
print(f"  merge kernel start -> last shard's flag seen (wait for the slowest) : {med(max(seen[e]) - R[e][1] for e in steps):8.1f}")
print(f"  flag arrival spread (last - first flag seen by the root)            : {med(max(seen[e]) - min(seen[e]) for e in steps):8.1f}")
print(f"  all flags seen -> merged, 'consumed' raised (reads over NVLink)     : {med(R[e][3] - R[e][2] for e in steps):8.1f}")
print(f"  merge kernel total                                                  : {med(R[e][3] - R[e][1] for e in steps):8.1f}")
print(f"  step period (merge done -> merge done of consecutive epochs)        : {med(R[b][3] - R[a][3] for a, b in zip(steps, steps[1:]) if b == a + 1):8.1f}")
if not instream:
    print(f"  root published -> root published of consecutive epochs              : {med(R[b][0] - R[a][0] for a, b in zip(steps, steps[1:]) if b == a + 1):8.1f}")
late = [max(range(len(seen[e])), key=lambda i: seen[e][i]) for e in steps]
print("  slowest shard histogram (rank: steps)                               :", {r: late.count(r) for r in sorted(set(late))})
