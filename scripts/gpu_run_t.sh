#!/bin/bash
# round 2, run T: where does the ADMM iteration (config 4) spend its time?
mkdir -p gpurun_out
for e in "" "PMX_NO_GRAPH=1"; do
env $e timeout 200 python bench.py --config 4 --steps 200 --warmup 5 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d.get('roofline') or {}; print('cfg4 [$e] it/s=%.1f ms=%.4f pass_ms=%.4f launches=%s' % (d['value'], d['ms_per_step'], r.get('avg_launch_ms'), d.get('gpu_launches')))
"
done 2>&1 | tee gpurun_out/r2t_cfg4.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2t_cfg4_launches.csv python bench.py --config 4 --steps 12 --warmup 3 --no-cpu > gpurun_out/r2t_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2t_cfg4_launches.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); ii=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[ii], r[ik][:40]),{})[r[im]]=r[iv]
for k,v in list(d.items())[-14:]:
    print(k, v)
PY
