"""Summarise an ncu --set full report: one row per launch with the metrics that explain a
memory-bound FP64 kernel (python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [out.csv])."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum']
stall = [h for h in hdr if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
idx = [hdr.index(w) for w in want if w in hdr]
out = [[hdr[i] for i in idx] + ['top_stalls'], [units[i] for i in idx] + ['']]
for r in rows[2:]:
    st = sorted(((float(r[hdr.index(h)] or 0), h.replace('smsp__pcsamp_warps_issue_stalled_', '')) for h in stall), reverse=True)[:4]
    out.append([r[i] for i in idx] + [' '.join('%s=%d' % (n, v) for v, n in st)])
w = csv.writer(open(sys.argv[2], 'w') if len(sys.argv) > 2 else sys.stdout)
w.writerows(out)
