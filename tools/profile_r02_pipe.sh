#!/bin/bash
# ncu captures of the two-pipeline step kernel (the kernel of bench.py's timed region at N=1), on the GPU box:
#   gpurun --timeout 1500 -- 'bash tools/profile_r02_pipe.sh'
O=gpurun_out
CMD="python bench.py --steps 10 --warmup 3 --no-extras --no-cpu"
$CMD > $O/r02_pipe_plain.json 2> $O/r02_pipe_plain.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/r02_launches_bench_pipe.csv $CMD > $O/r02_launches_bench_pipe.log 2>&1
echo "launch list rc=$?"
# the timed launch: 10 steps of k_step_pipe (launches: 6 x 50 pretraining steps, 3 warm-up steps, then the timed 10)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_pipe -s 7 -c 1 -o $O/r02_step_pipe $CMD > $O/r02_ncu_pipe.log 2>&1
echo "pipe rc=$?"
ncu -i $O/r02_step_pipe.ncu-rep --page raw --csv > $O/r02_step_pipe.raw.csv 2> /dev/null
ncu -i $O/r02_step_pipe.ncu-rep --page source --csv > $O/r02_step_pipe.source.csv 2> /dev/null
python - <<PY
import csv, collections
rows=list(csv.reader(open("$O/r02_step_pipe.source.csv")))
start=[i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"]
seg=rows[start[0]+1:]
hdr=seg[0]; data=seg[1:]
ix={n:i for i,n in enumerate(hdr)}
stalls=[n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot=collections.Counter()
for r in data:
    if len(r)<len(hdr): continue
    for n in stalls: tot[n]+=int(r[ix[n]] or 0)
T=sum(tot.values())
open("$O/r02_step_pipe_stalls.txt","w").write("\n".join(f"{n} {v} {100*v/T:.1f}%" for n,v in tot.most_common(10)))
PY
rm -f $O/r02_step_pipe.source.csv
ls -la $O | tail -12
