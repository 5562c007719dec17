"""Summarise an .ncu-rep (run where ncu is installed): key metrics + hot code runs.  usage: ncu_summary.py rep [min_share]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ['gpu__time_duration.sum', 'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum', 'pipe_fma.avg.pct_of_peak_sustained_active', 'pipe_alu.avg.pct_of_peak_sustained_active',
        'pipe_lsu.avg.pct_of_peak_sustained_active', 'pipe_xu.avg.pct_of_peak_sustained_active', 'pipe_fp64.avg.pct_of_peak_sustained_active', 'bank_conflicts_pipe_lsu_mem_shared.sum',
        'thread_inst_executed_per_inst', 'smsp__pcsamp_warps_issue_stalled', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__warps_active.avg.per_cycle',
        'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed_op_shared_atom.sum', 'sm__throughput.avg.pct', 'launch__occupancy_limit']
for k in sorted(d):
    if any(s in k for s in keys):
        v, u = d[k]
        if 'not_issued' in k or 'per_second' in k or 'peak_sustained' in k and 'pct' not in k: continue
        try:
            if 'pcsamp' in k and float(v) < 20000: continue
        except ValueError: pass
        print(f"{k:90s} {v:>20s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ia, isrc, isamp, iex, ithr = (hdr.index(x) for x in ("Address", "Source", "# Samples", "Instructions Executed", "Avg. Threads Executed"))
tot = sum(int(r[iex]) for r in data); ts = sum(int(r[isamp]) for r in data)
print(f"kernel: {rows[0][1][:120]}\ntotal warp-instructions {tot:.4g}, samples {ts}")
runs = []; cur = None
for k, r in enumerate(data):
    ex = int(r[iex]); key = round(ex / max(tot / 3000.0, 1))
    if cur and abs(cur[2] - key) <= max(1, 0.03 * key): cur[1] = k; cur[3] += ex; cur[4] += int(r[isamp])
    else: cur = [k, k, key, ex, int(r[isamp])]; runs.append(cur)
for a, b, key, ex, s in runs:
    if ex / tot > min_share:
        ops = {}
        for r in data[a:b + 1]:
            t = r[isrc].split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]; ops[op] = ops.get(op, 0) + 1
        print(f"lines {a}-{b} ({b - a + 1} instr) exec/instr {ex / (b - a + 1):.3g} inst_share {ex / tot:.3f} sample_share {s / ts:.3f} thr {data[a][ithr]} {dict(sorted(ops.items(), key=lambda x: -x[1]))}")
