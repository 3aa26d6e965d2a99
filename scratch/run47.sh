python tools/profile_tc.py simple_spread 24 2048 1024 1 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_round5.csv python tools/profile_tc.py simple_spread 24 2048 1024 1 > gpurun_out/ncu_c.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_round5.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mn=h.index('Metric Name'); mv=h.index('Metric Value'); idc=h.index('ID')
d={}
for r in rows[hi+1:]:
    if len(r)>mv: d.setdefault((r[idc], r[kn][:44]),{})[r[mn]]=r[mv]
for (i,k),v in list(d.items())[-9:]: print(i, k, v.get('gpu__time_duration.sum'), v.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'))
PY
