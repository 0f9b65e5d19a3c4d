python profiles/microbench.py 2>&1 | head -5 | tee gpurun_out/micro.log
python profiles/ncu_variants.py > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"step_pool_kernel|step_obs_kernel|encode_kernel" -s 29 -c 4 --csv --log-file gpurun_out/light.csv python profiles/ncu_variants.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/light.csv')) if len(r)>10]
h=rows[0]; 
for r in rows[1:]:
    print(r[h.index('ID')], r[h.index('Kernel Name')][:45], r[h.index('Metric Name')], r[h.index('Metric Value')])
PY
