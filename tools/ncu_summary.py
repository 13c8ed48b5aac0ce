#!/usr/bin/env python
"""Summarise an `ncu --set full` report: python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.json]
Reads the raw page (`ncu -i ... --page raw --csv`) and keeps the metrics the roofline discussion uses."""
import csv
import io
import json
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__cluster_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tensor.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg.per_second', 'smsp__inst_executed.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum', 'l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed']


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for r in data:
        d = {'kernel': r[hdr.index('Kernel Name')].split('(')[0], 'grid': r[hdr.index('Grid Size')], 'block': r[hdr.index('Block Size')]}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = ('%s %s' % (r[i], units[i])).strip()
        out.append(d)
    text = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(text + '\n')
    print(text)


if __name__ == '__main__':
    main()
