#!/bin/bash
# Round 2, twenty-fourth GPU pass (1 GPU): launch lists of the k = 11 and k = 31 steps.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1
export HULK_B200_FEEDER=0
for k in 11 31; do
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'k[0123]_' --launch-skip 40 -c 60 \
    --csv --log-file gpurun_out/r02p_k${k}_launches.csv python bench.py --k $k --s 128 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02p_k${k}.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/r02p_k${k}_launches.csv") if l.startswith('"'))]
h=rows[0]; ix={x:i for i,x in enumerate(h)}
agg={}
for r in rows[1:]:
    n=r[ix["Kernel Name"]].split("(")[0].replace("void ","")
    m=r[ix["Metric Name"]]; v=float(r[ix["Metric Value"]])
    a=agg.setdefault(n,{"n":0,"ns":0,"inst":0})
    if m=="gpu__time_duration.sum": a["n"]+=1; a["ns"]+=v
    else: a["inst"]+=v
print("k=${k}")
for n,a in agg.items():
    if a["n"]: print("  %-44s %3d launches  %8.2f us avg  %10.0f warp instr avg"%(n,a["n"],a["ns"]/a["n"]/1e3,a["inst"]/a["n"]))
PY
done
