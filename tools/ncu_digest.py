"""Digest an .ncu-rep of the fused kernel: headline metrics, stall mix, barrier waits, softmax loop mix."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
        'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active']
for h, u, v in zip(hdr, units, vals):
    if h in keep:
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter()
total = 0
recs = []
for r in data:
    try:
        ns = int(r[ix['# Samples']])
    except Exception:
        continue
    total += ns
    d = {s: int(r[ix[s]] or 0) for s in stalls}
    for s in stalls:
        tot[s] += d[s]
    recs.append((ns, r[ix['Source']], d, int(r[ix['Instructions Executed']] or 0)))
print("total samples", total)
print("stall mix:", ", ".join(f"{s[6:]} {100*v/total:.1f}%" for s, v in tot.most_common(8)))
print("top lines:")
for ns, s, d, ie in sorted(recs, key=lambda x: -x[0])[:10]:
    print(f"  {ns:6d} {100*ns/total:5.1f}% ie={ie:8d} {s[:60]:60s} {sorted(d.items(), key=lambda x: -x[1])[:2]}")
names = {0x30000: 'q_full0', 0x30008: 'q_full1', 0x30010: 'k_full', 0x30018: 'k_full1', 0x30020: 'k_empty', 0x30028: 'k_empty1',
         0x30030: 'v_full', 0x30038: 'v_full1', 0x30040: 'v_empty', 0x30048: 'v_empty1', 0x30050: 's_full', 0x30058: 's_full1',
         0x30060: 'p_ready0', 0x30068: 'p_ready1', 0x30070: 'o_done', 0x30078: 'o_done1'}
print("barrier waits (samples at TRYWAIT + following BRA):")
for i, r in enumerate(data[:-1]):
    s = r[ix['Source']]
    if 'TRYWAIT' in s:
        m = re.search(r'0x3[0-9a-f]{4}', s)
        nm = names.get(int(m.group(0), 16), '?') if m else '(reg)'
        ie = int(r[ix['Instructions Executed']] or 0)
        ns = int(r[ix['# Samples']] or 0) + int(data[i + 1][ix['# Samples']] or 0)
        if ie > 0 and ns > 5:
            print(f"  {nm:9s} spins={ie:8d} samples={ns}")
ies = collections.Counter(r[3] for r in recs if r[3] > 100000)
if ies:
    loop_ie = max(ies, key=lambda k: ies[k])
    cnt, samp = collections.Counter(), collections.Counter()
    n = 0
    for ns, s, d, ie in recs:
        if abs(ie - loop_ie) <= loop_ie * 0.03:
            op = [o for o in s.split() if not o.startswith('@')][0].split('.')[0]
            cnt[op] += 1
            samp[op] += ns
            n += 1
    print(f"softmax loop: ie={loop_ie}, {n} instr/warp/tile:", {k: (v, samp[k]) for k, v in cnt.most_common(12)})
