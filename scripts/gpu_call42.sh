#!/bin/bash
# in-situ L2 behaviour of the Krylov kernels: warm caches (no flush between kernels), one pass of three metrics
set -u
O=gpurun_out/c42; mkdir -p $O
CMD="python bench.py --horizon 10 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --no-slab --no-ensemble --profile-steps 5"
VCH_NO_GRAPHS=1 timeout 420 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -k 'regex:rows16|cols16' -s 300 -c 120 --csv --log-file $O/insitu.csv $CMD > $O/ncu.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, re, collections
rows=[l for l in open('gpurun_out/c42/insitu.csv') if not l.startswith('==')]
rd=csv.DictReader(rows)
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rd:
    name=re.sub(r"^void |\(.*$","",r["Kernel Name"]).replace("vch::","")
    try: v=float(r["Metric Value"].replace(",",""))
    except: continue
    u=r["Metric Unit"]
    if r["Metric Name"].startswith("dram__bytes"): v*= {"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}.get(u,1)
    if r["Metric Name"].startswith("gpu__time"): v*= {"ns":1e-3,"us":1,"ms":1e3}.get(u,1e-3)
    agg[name][r["Metric Name"]].append(v)
for n,m in agg.items():
    f=lambda k: sum(m[k])/max(1,len(m[k]))
    print(f"{n:34s} n={len(m['gpu__time_duration.sum']):3d} time {f('gpu__time_duration.sum'):7.2f} us  dram read {f('dram__bytes_read.sum')/1e6:7.2f} MB  write {f('dram__bytes_write.sum')/1e6:6.2f} MB  L2 hit {f('lts__t_sector_hit_rate.pct'):5.1f} %")
PY
