#!/bin/bash
# Developer A/B: ring shape of k_dense_decode_tma for small scans (crowd), kernel alone (graph-timed) and inside the step.
Q="--workload crowd --regions 3 --skip-cpu-baseline --skip-gpu-eager --skip-half-maps --skip-mlp --skip-train-tail --skip-e2e"
for cfg in "- - -" "64 2 3" "32 2 4" "64 2 2" "- - -" "64 2 3"; do
    set -- $cfg
    if [ "$1" = "-" ]; then unset SIHL_DECODE_ROWS SIHL_DECODE_STAGES SIHL_DECODE_CTAS_PER_SM; else export SIHL_DECODE_ROWS=$1 SIHL_DECODE_STAGES=$2 SIHL_DECODE_CTAS_PER_SM=$3; fi
    python bench.py $Q 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.readlines()[-1])
print('$cfg', 'step %.2f us (min %.2f)  cand-first %.2f us  kernel %.2f us frac %.3f  step frac %.3f' % (d['ms_per_step'] * 1e3, d['timing']['ms_per_step_min'] * 1e3, d['other_decode_mode']['ms_per_step'] * 1e3, d['roofline']['kernel_ms'] * 1e3, d['roofline']['frac'], d['roofline_step']['frac']))
"
done
