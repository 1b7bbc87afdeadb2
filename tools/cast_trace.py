"""Developer aid: where in time do the CTAs of the background cast (cast-ahead) run relative to the fused kernel they share
the SMs with?  Queues a few c3 passes with SDPA_CAST_TRACE set and digests the per-CTA %globaltimer stamps of the LAST
background cast against the begin/end stamps of the last two fused kernels.  Usage: python tools/cast_trace.py [out]"""
import os
import statistics as st
import sys
from pathlib import Path

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/cast_trace.txt"
if os.environ.get("NO_TRACE") != "1":
    os.environ["SDPA_CAST_TRACE"] = out
os.environ.setdefault("SDPA_STAGE_TIMING", "0")
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import sdpa_b200  # noqa: E402

m, n = 8192, 65536
g = torch.Generator().manual_seed(0)
Q = torch.randn(m, 128, dtype=torch.float64, generator=g).cuda()
K = torch.randn(n, 128, dtype=torch.float64, generator=g).cuda()
V = torch.randn(n, 128, dtype=torch.float64, generator=g).cuda()
R = torch.zeros(m, 128, dtype=torch.float64, device="cuda")
import time  # noqa: E402
passes = int(os.environ.get("PASSES", "40"))
with sdpa_b200.Context(precision="bf16") as ctx:
    for _ in range(4):
        ctx.attention_device_full([K.data_ptr()], [V.data_ptr()], [n], 128, 128, [Q.data_ptr()], R.data_ptr(), m, blocking=False)
    ctx.synchronize()
    kd, vd, qd = [K.data_ptr()], [V.data_ptr()], [Q.data_ptr()]
    t0 = time.perf_counter()
    for _ in range(passes):
        ctx.attention_device_full(kd, vd, [n], 128, 128, qd, R.data_ptr(), m, blocking=False)
    t1 = time.perf_counter()
    ctx.synchronize()
    t2 = time.perf_counter()
    print(f"{passes} queued passes: host issue {(t1 - t0) * 1e6 / passes:.1f} us per pass, wall incl. wait {(t2 - t0) * 1e6 / passes:.1f} us per pass")
passes += 4
if os.environ.get("NO_TRACE") == "1":
    sys.exit(0)
lines = open(out).read().splitlines()
f = list(map(int, lines[0].split()[1:5]))
last = (passes - 1) & 1                      # parity slot of the last pass P; the last cast ran beside fused(P-1)
fb, fe = f[2 * (1 - last)], f[2 * (1 - last) + 1]      # fused(P-1)
nb, ne = f[2 * last], f[2 * last + 1]                   # fused(P)
ctas = [tuple(map(int, l.split()[1:4])) for l in lines[1:]]
starts = [(s - fb) / 1e3 for _, s, _ in ctas]
ends = [(e - fb) / 1e3 for _, _, e in ctas]
durs = [(e - s) / 1e3 for _, s, e in ctas]
print(f"fused(P-1): 0 .. {(fe - fb) / 1e3:.1f} us;  fused(P): {(nb - fb) / 1e3:.1f} .. {(ne - fb) / 1e3:.1f} us  (all relative to fused(P-1) begin)")
print(f"background cast, {len(ctas)} CTAs: start min/median/max {min(starts):.1f} / {st.median(starts):.1f} / {max(starts):.1f} us, "
      f"end min/median/max {min(ends):.1f} / {st.median(ends):.1f} / {max(ends):.1f} us, duration median {st.median(durs):.1f} us")
