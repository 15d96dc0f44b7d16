#!/bin/bash
mkdir -p gpurun_out
python tools/prof_msm.py 18 3 > gpurun_out/r02_prof_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum --clock-control none --launch-skip 13 -c 12 --csv --log-file gpurun_out/r02_msm_launches.csv python tools/prof_msm.py 18 3 > gpurun_out/r02_ncu_list.log 2>&1
tail -3 gpurun_out/r02_ncu_list.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_msm_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault((r[ii], r[ki][:40]), {})[r[mi]]=r[vi]
for (i,k),m in list(cur.items())[-14:]:
    print(i, k, {a.split('__')[-1][:28]:b for a,b in m.items()})
PY
