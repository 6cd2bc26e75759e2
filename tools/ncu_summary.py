"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into the handful of metrics DESIGN.md / profiles/ cite.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [substring ...]
"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active',
        'sm__inst_executed_pipe_tensor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit', 'smsp__issue_active.avg.pct',
        'sm__throughput.avg.pct', 'smsp__inst_executed.sum', 'sm__pipe_fma_cycles_active.avg.pct',
        'sm__pipe_fmaheavy_cycles_active.avg.pct', 'sm__pipe_alu_cycles_active.avg.pct',
        'sm__inst_executed_pipe_lsu', 'sm__inst_executed_pipe_xu.avg.pct', 'lts__t_bytes.sum ',
        'lts__throughput.avg.pct', 'l1tex__throughput.avg.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ',
        'smsp__average_warps_issue_stalled', 'smsp__average_warp_latency_issue_stalled',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__cycles_elapsed.avg ', 'smsp__cycles_active.avg ', 'sm__pipe_shared_cycles_active',
        'smsp__inst_executed_op_shared', 'sm__inst_executed_pipe_uniform', 'smsp__thread_inst_executed_per_inst']


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('== kernel:', r[hdr.index('Kernel Name')], '| grid', r[hdr.index('Grid Size')], '| block',
              r[hdr.index('Block Size')])
        for h, u, v in zip(hdr, units, r):
            hh = h + ' '
            if any(k in hh for k in KEYS) or any(e in h for e in extra):
                print('  %-90s %-14s %s' % (h, u, v))


if __name__ == '__main__':
    main()
