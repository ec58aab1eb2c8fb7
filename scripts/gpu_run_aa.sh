#!/bin/bash
# round 2, run AA: launch list of an adaprox iteration (config 3, one GPU)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2aa_cfg3_launches.csv python bench.py --config 3 --steps 4 --warmup 2 --no-cpu > gpurun_out/r2aa_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2aa_cfg3_launches.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); ii=hdr.index('ID')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ik].split('(')[0][-40:]),{})[r[im]]=float(r[iv].replace(',',''))
items=list(d.items())
# print the last ~45 launches in order
for (i,k),v in items[-48:]:
    print(i, '%-42s %8.1f us  rd %7.1f MB wr %7.1f MB' % (k, v.get('gpu__time_duration.sum',0)/1e3 if v.get('gpu__time_duration.sum',0)>1e3 else v.get('gpu__time_duration.sum',0), v.get('dram__bytes_read.sum',0)/1e6, v.get('dram__bytes_write.sum',0)/1e6))
PY
