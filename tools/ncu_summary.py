"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
print("kernel", [d[idx['Kernel Name']].replace('void ', '')[:22] for d in data])
for w in want:
    if w in idx:
        print(f"{w} [{units[idx[w]]}]", [d[idx[w]][:12] for d in data])
ks = [h for h in hdr if 'issue_stalled' in h and 'pcsamp' in h and 'not_issued' not in h]
for k, d in enumerate(data):
    tot = sum(float(d[idx[h]] or 0) for h in ks)
    if tot == 0:
        continue
    top = sorted(ks, key=lambda h: -float(d[idx[h]] or 0))[:7]
    print(d[idx['Kernel Name']].replace('void ', '')[:22], "stalls:",
          ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * float(d[idx[h]]) / tot:.1f}%" for h in top))
