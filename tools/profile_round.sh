#!/bin/bash
# Developer tool (run under gpurun, 1 GPU): the ncu evidence committed under profiles/ for one round.
#   tools/profile_round.sh r01   ->  gpurun_out/<tag>_launches.csv, <tag>_full.ncu-rep, <tag>_cf.ncu-rep
# Every command runs plainly first (exit code checked) and only then under ncu; numbers printed under ncu are not used.
tag=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 12 --warmup 3 --regions 1 --no-graph --lanes 1 --skip-e2e --skip-cpu-baseline --skip-gpu-eager --skip-candidate-first --skip-train-tail --skip-half-maps"
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
# 1. every launch of our kernels with its device time (cold cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 200 --csv --log-file gpurun_out/${tag}_launches.csv \
    $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "launch list exit $?"
# 2. full set + source counters for one whole step (5 kernels) after 3 warm-up steps
ncu --set full --clock-control none --import-source on -k regex:^k_ -s 15 -c 5 -o gpurun_out/${tag}_full \
    $CMD > gpurun_out/${tag}_ncu_full.log 2>&1; echo "full exit $?"
# 3. the candidate-first decode kernel
CF="$CMD --decode-mode candidate_first"
$CF > gpurun_out/${tag}_plain_cf.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_candidate -s 3 -c 1 -o gpurun_out/${tag}_cf \
    $CF > gpurun_out/${tag}_ncu_cf.log 2>&1; echo "cf exit $?"
