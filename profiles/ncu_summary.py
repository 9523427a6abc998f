"""Condense an `ncu --set full` report into one CSV row per captured launch:
    ncu -i prof.ncu-rep --page raw --csv > raw.csv ; python profiles/ncu_summary.py raw.csv > summary.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'smsp__cycles_active.avg', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor']
idx = [(w, hdr.index(w)) for w in want if w in hdr]
out = csv.writer(sys.stdout)
out.writerow(['%s [%s]' % (w, units[i]) if units[i] else w for w, i in idx])
for r in rows[2:]:
    out.writerow([r[i][:60] for _, i in idx])
