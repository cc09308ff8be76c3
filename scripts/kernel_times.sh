#!/bin/bash
# usage: scripts/kernel_times.sh <tag> <bench args...> : per-kernel time + DRAM bytes of one warmed-up bench step (ncu, serialised)
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s ${SKIP:-30} -c ${COUNT:-60} --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --reads-per-step 1000000 --no-cpu-baseline --no-e2e "$@" > gpurun_out/launches_$tag.log 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/launches_$tag.csv")))
for i,r in enumerate(rows):
    if "Kernel Name" in r: hdr=r; start=i+1; break
ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit"); mi=hdr.index("Metric Name")
agg=collections.defaultdict(lambda: collections.defaultdict(float)); cnt=collections.Counter()
for r in rows[start:]:
    if len(r)<=vi: continue
    n=r[ki].split("(")[0][:34]; v=float(r[vi].replace(",","")); u=r[ui]; m=r[mi]
    scale={"us":1e-3,"ns":1e-6,"ms":1,"byte":1e-9,"Kbyte":1e-6,"Mbyte":1e-3,"Gbyte":1}.get(u,1)
    agg[n][m]+=v*scale
    if m=="gpu__time_duration.sum": cnt[n]+=1
for n,d in sorted(agg.items(), key=lambda x:-x[1]["gpu__time_duration.sum"]):
    print("%-36s x%3d %9.2f ms/launch  read %7.2f GB write %7.2f GB per launch" % (n,cnt[n],d["gpu__time_duration.sum"]/cnt[n],d["dram__bytes_read.sum"]/cnt[n],d["dram__bytes_write.sum"]/cnt[n]))
PY
