"""Brief of one .ncu-rep: headline metrics, stall reasons, and the hot SASS lines.  usage: ncu_brief.py rep [min_share]"""
import csv, subprocess, sys
rep = sys.argv[1]
share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h, v = r[0], r[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__throughput.avg.pct', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__warps_active.avg.pct', 'dram__throughput.avg.pct', 'l1tex__throughput.avg.pct', 'launch__registers_per_thread', 'lts__t_bytes.sum ',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__inst_executed_pipe_lsu', 'smsp__inst_executed_pipe']
for i, n in enumerate(h):
    if any(n.startswith(w.strip()) for w in want) and 'per_second' not in n:
        print(f"{n:80s} {r[1][i]:8s} {v[i]}")
st = [(float(v[i]), n) for i, n in enumerate(h) if n.startswith('smsp__average_warps_issue_stalled') and n.endswith('_per_issue_active.ratio') or
      (n.startswith('smsp__average_warp_latency_issue_stalled') and n.endswith('.ratio'))]
for a, n in sorted(st, reverse=True)[:8]:
    print(f"  stall {n.replace('smsp__average_warps_issue_stalled_','').replace('smsp__average_warp_latency_issue_stalled_','')[:40]:42s} {a:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
r = list(csv.reader(src.splitlines()))
i0 = next(i for i, x in enumerate(r) if len(x) > 5)
h = r[i0]
ie = h.index('Instructions Executed'); sc = h.index('Source'); ss = h.index('Warp Stall Sampling (All Samples)')
rows = [(int(x[ie] or 0), int(x[ss] or 0), x[sc]) for x in r[i0 + 1:] if len(x) > ie]
tot = sum(a for a, _, _ in rows); tots = sum(b for _, b, _ in rows)
print("instructions", tot, "samples", tots)
for idx, (a, b, s_) in enumerate(rows):
    if a > tot * share or b > tots * share * 2:
        print(f"{idx:5d} {a:10d} {100*b/max(tots,1):5.1f}%  {s_[:100]}")
