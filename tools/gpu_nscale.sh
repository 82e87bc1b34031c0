#!/bin/bash
# ncu device time of a workload's kernels against the env count (cold cache, serialised): tools/gpu_nscale.sh <workload> <tag> [ENV=1]
O=gpurun_out
for n in 1024 2048 4096 8192 16384 32768; do
  env ${3:-_X=1} ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/ns_$2_$n.csv python bench.py --workload $1 --envs $n --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --only --no-graph > /dev/null 2>&1
  python - $O/ns_$2_$n.csv $n <<'P'
import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
d=collections.defaultdict(list)
for r in rows[1:]:
    d[r[ik].split('(')[0][:40]].append(float(r[iv].replace(',','')))
print(sys.argv[2], {k:round(sorted(v)[len(v)//2]/1000 if max(v)>1000 else sorted(v)[len(v)//2],2) for k,v in d.items() if 'synth' not in k and 'Fill' not in k})
P
done
