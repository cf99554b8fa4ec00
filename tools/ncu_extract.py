#!/usr/bin/env python
"""Extract the metrics quoted in profiles/README.md from an `ncu --set full` report:
    python tools/ncu_extract.py report.ncu-rep out.csv
One row per (launch, metric): duration, DRAM bytes, pipe utilisation, registers, shared memory and the warp stall reasons."""
import csv
import subprocess
import sys

KEEP = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum')


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as f:
        w = csv.writer(f)
        w.writerow(['launch', 'kernel', 'metric', 'unit', 'value'])
        for i, r in enumerate(rows[2:]):
            name = r[hdr.index('Kernel Name')]
            for j, h in enumerate(hdr):
                if h in KEEP or ('warp_issue_stalled' in h and h.endswith('per_warp_active.pct') and float(r[j].replace(',', '') or 0) >= 2.0):
                    w.writerow([i, name[:60], h, units[j], r[j]])


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
