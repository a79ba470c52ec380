"""Digest an `ncu --page raw --csv` dump: the handful of metrics that decide what bounds each kernel."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'sm__cycles_elapsed.max']
stall = [h for h in hdr if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('----', d['Kernel Name'][:60])
    for w in want:
        if w in d:
            print(f"  {w:75s} {d[w]:>16s} {units[hdr.index(w)]}")
    st = sorted(((float(d[h].replace(',', '')), h) for h in stall if d[h] not in ('', 'n/a')), reverse=True)[:6]
    for v, h in st:
        print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', '').replace('_per_issue_active.ratio', ''):60s} {v:8.2f}")
