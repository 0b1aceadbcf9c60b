timeout 600 ncu --set full --clock-control none --import-source on -k regex:vx_fill_lattice -s 2 -c 1 -o gpurun_out/r02_fill python bench.py --region 256 --steps 1 --warmup 3 --device-only > /dev/null 2>&1
ncu -i gpurun_out/r02_fill.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    for k,v in zip(h,r):
        if 'issue_stalled' in k and 'ratio' in k and 'not_issued' not in k: print(k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), v)
        if k in ('gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed.sum','sm__icc_request_hit_rate.pct','smsp__inst_executed.sum','launch__registers_per_thread','launch__grid_size'): print(k, v)
"
