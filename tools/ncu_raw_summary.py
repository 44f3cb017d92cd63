#!/usr/bin/env python
"""One CSV line per profiled kernel of an ncu report (time, DRAM bytes, instructions, lane occupancy, issue activity, L2 hit
rate ...): the summaries committed under profiles/.

    python tools/ncu_raw_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.csv
"""
import csv,sys,subprocess,io
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
r=list(csv.reader(io.StringIO(out)))
h=r[0]
want=[w for w in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','lts__t_sector_hit_rate.pct','launch__grid_size','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','launch__block_size'] if w in h]
print('kernel,'+','.join(want)); print('unit,'+','.join(r[1][h.index(w)] for w in want))
for row in r[2:]:
    print(row[h.index('Kernel Name')].split('(')[0]+','+','.join(row[h.index(w)] for w in want))
