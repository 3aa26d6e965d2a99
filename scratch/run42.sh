python tools/profile_tc_critic.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_tcc.csv python tools/profile_tc_critic.py > gpurun_out/ncu_c.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_tcc.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit')
for r in rows[hi+1:]:
    if len(r)>mv: print(r[kn][:60], r[mv], r[mu])
PY
