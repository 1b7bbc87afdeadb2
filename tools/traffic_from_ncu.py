"""profiles/fused_kernel_traffic.json from `ncu --set full` captures: DRAM bytes per launch of each fused kernel, keyed by the
kernel name sdpa_last_kernel() reports.  bench.py reads that file for roofline.traffic (a STATIC number from the capture named
in `source`, not a live measurement).  Usage: python tools/traffic_from_ncu.py name=path.ncu-rep [name=path ...]"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
out_path = ROOT / "profiles" / "fused_kernel_traffic.json"
table = json.loads(out_path.read_text()) if out_path.exists() else {}
for arg in sys.argv[1:]:
    name, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    best = None
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))

        def val(key):
            x = float(d[key].replace(",", ""))
            unit = u[key].lower()
            return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        rec = {"dram_bytes_per_launch": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
               "dram_read": int(val("dram__bytes_read.sum")), "dram_write": int(val("dram__bytes_write.sum")),
               "kernel": d.get("Kernel Name", "?")[:80], "grid": d.get("launch__grid_size"), "duration_us_under_ncu": d.get("gpu__time_duration.sum"),
               "tensor_pipe_active_pct": d.get("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed") or d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
               "source": f"ncu --set full --clock-control none, {Path(rep).name} (committed digest under profiles/r02/)"}
        if best is None or float(rec["duration_us_under_ncu"].replace(",", "")) > float(best["duration_us_under_ncu"].replace(",", "")):
            best = rec
    if best:
        table[name] = best
        print(name, best)
out_path.write_text(json.dumps(table, indent=1) + "\n")
