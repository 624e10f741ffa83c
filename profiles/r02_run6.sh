# round 2, sixth GPU pass: tests; the large-cloud kernels one by one (launch list + ncu --set full of the search and
# Mahalanobis kernels of update_correspondences and of linearize at 20 M points); pool after reverting the prefetch
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_tests6.txt
cat gpurun_out/r02_tests6.txt
timeout 300 python profiles/pool_probe.py --no-launch-rate --streams 128 > gpurun_out/r02_probe6.txt 2>&1
cut -c1-400 gpurun_out/r02_probe6.txt
C4="python bench.py --workload c4 --steps 3 --roofline-reps 3 --no-cpu-baseline"
timeout 600 $C4 > gpurun_out/r02_c4_6.json 2>/dev/null; cut -c1-1500 gpurun_out/r02_c4_6.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c4.csv $C4 > /dev/null 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"corr_search|maha_kernel|linearize_kernel" -s 12 -c 4 -o gpurun_out/prof_c4_r02 -f $C4 > gpurun_out/ncu_c4_r02.log 2>&1
echo ncu rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_c4.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg={}
for r in rows[1:]:
    k=r[ki].split('(')[0][-60:]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[vi].replace(',',''))
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print(f"{k:62s} n={n:4d} total={t/1e6:9.3f} ms mean={t/n/1e3:9.1f} us")
PY
