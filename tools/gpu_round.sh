#!/bin/bash
# One GPU-box pass: parity tests, smoke, every bench line, then the ncu launch list and one full capture.
set -u
O=gpurun_out
TAG=${1:-r1}
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke_$TAG.log
python bench.py > $O/bench_${TAG}_default.json 2> $O/bench_${TAG}_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_${TAG}_reference.json 2> $O/bench_${TAG}_reference.err; echo "ref rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --only"
$CMD > $O/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1
echo "ncu launches rc=$?"
$CMD > $O/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_ingest_gray_std|k_ingest_atari_tma|k_observe_peripheral' -s 8 -c 4 -f -o $O/prof_$TAG $CMD > $O/ncu_f_$TAG.log 2>&1
echo "ncu full rc=$?"
for pair in "atari_fixed:k_ingest_atari_tma:rgb" "atari_flexible:k_observe_flexible_v3:flex" "dmc_fixed:k_ingest_dmc|k_observe_fixed_crop:dmc"; do
  w=${pair%%:*}; rest=${pair#*:}; k=${rest%%:*}; t=${rest#*:}
  C2="python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --only"
  $C2 > $O/plain_${TAG}_$t.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s 8 -c 2 -f -o $O/prof_${TAG}_$t $C2 > $O/ncu_f_${TAG}_$t.log 2>&1
  echo "ncu $t rc=$?"
done
tail -c 600 $O/bench_${TAG}_default.json
