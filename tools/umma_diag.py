"""GPU diagnostic for the bf16 tcgen05 kernel: runs a ladder of shapes (each in its own
subprocess with a timeout, so a trapped kernel cannot take the session down) and prints the
error against the oracle, with a per-block error map when a case fails.
Usage (on the GPU box):  python tools/umma_diag.py [--out gpurun_out/umma_diag.txt]"""
import argparse
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

CASES = [
    # m, n, splits, q_batch
    (128, 128, 1, 0), (256, 128, 1, 0), (128, 256, 1, 0), (256, 512, 1, 0), (256, 512, 2, 0),
    (100, 200, 1, 0), (300, 1000, 0, 0), (1024, 4096, 0, 0), (1024, 4096, 0, 256), (4096, 4096, 0, 0),
]


def child(m, n, splits, q_batch):
    import numpy as np
    import sdpa_b200
    from oracle import oracle
    Q, K, V = oracle.make_inputs(m, n, 128, 128, seed=m * 7 + n)
    ref = oracle.attention_f64_numpy(Q, K, V)
    # what the kernel should compute exactly: bf16-rounded operands, fp64 math
    Qb, Kb, Vb = (oracle.bf16_round(a).astype(np.float64) for a in (Q, K, V))
    ref_b = oracle.attention_f64_numpy(Qb, Kb, Vb)
    with sdpa_b200.Context(precision="bf16", kv_splits=splits, q_batch=q_batch) as ctx:
        ctx.load_kv_host_full(K, V)
        got = ctx.attention_host(Q)
        tm = ctx.last_timings()
        kern = ctx.last_kernel()
    err = np.abs(got - ref)
    err_b = np.abs(got - ref_b)
    res = dict(m=m, n=n, splits=splits, q_batch=q_batch, kernel=kern, max_err=float(err.max()),
               max_err_vs_bf16_inputs=float(err_b.max()), nan=int(np.isnan(got).sum()), fused_ms=tm["fused_ms"],
               total_ms=tm["total_ms"])
    if not (err.max() < 1e-2):
        rb = -(-m // 32)
        blocks = [[float(np.nanmax(err[r * 32:(r + 1) * 32, c * 32:(c + 1) * 32])) for c in range(4)] for r in range(min(rb, 16))]
        res["block_err_32x32"] = blocks
        res["got_row0"] = [float(x) for x in got[0, :8]]
        res["ref_row0"] = [float(x) for x in ref[0, :8]]
        res["mean_abs_got"] = float(np.nanmean(np.abs(got)))
        res["mean_abs_ref"] = float(np.mean(np.abs(ref)))
    print("RESULT " + json.dumps(res))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", nargs=4, type=int)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.child:
        child(*a.child)
        return
    lines = []
    for case in CASES:
        cmd = [sys.executable, __file__, "--child"] + [str(x) for x in case]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
            out = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            line = out[0] if out else f"FAIL case={case} rc={r.returncode} stdout={r.stdout[-600:]!r} stderr={r.stderr[-900:]!r}"
        except subprocess.TimeoutExpired:
            line = f"TIMEOUT case={case}"
        print(line, flush=True)
        lines.append(line)
    if a.out:
        Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        Path(a.out).write_text("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
