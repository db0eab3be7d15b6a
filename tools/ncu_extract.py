#!/usr/bin/env python3
"""tools/ncu_extract.py raw.csv [pattern...] -- print selected metrics per kernel from `ncu -i rep --page raw --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pats = sys.argv[2:] or ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__occupancy_limit', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'lts__throughput.avg.pct', 'l1tex__throughput.avg.pct', 'smsp__inst_executed.sum', 'issue_active',
    'smsp__average_warp', 'smsp__warp_issue_stalled', 'shared_mem', 'dram__cycles_active']
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:60], 'id', r[0])
    for i, h in enumerate(hdr):
        if any(h == p or (p in h) for p in pats):
            if r[i] not in ('', '0', 'n/a'):
                print("  %-90s %s %s" % (h, r[i], units[i]))
