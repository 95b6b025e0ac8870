#!/bin/bash
# Developer A/B of the end-to-end leg (host-resident class / box maps read in place over PCIe):
#   SIHL_HOST_ROWS = 0 (device-path lane layout), 1 (8 lanes x 16 B), 2 (whole row per warp instruction) for the positive-row
#   kernel; SIHL_HOST_CAND = 0 / 1: device-path / whole-row reads in the candidate kernel; x steps in flight.
# Usage (GPU box): bash tools/e2e_ab.sh "2 1 3" "2 0 3" ... > gpurun_out/e2e_ab.txt     (rows, cand, lanes)
common="--steps 20 --warmup 5 --regions 1 --skip-cpu-baseline --skip-gpu-eager --skip-half-maps --skip-mlp --skip-train-tail --skip-candidate-first --e2e-steps 120"
[ $# -eq 0 ] && set -- "2 1 3" "2 0 3" "1 0 3" "0 0 3" "2 1 3" "2 0 3"
for cfg in "$@"; do
    set -- $cfg
    echo "== SIHL_HOST_ROWS=$1 SIHL_HOST_CAND=$2 e2e-lanes=$3"
    SIHL_HOST_ROWS=$1 SIHL_HOST_CAND=$2 python bench.py $common --e2e-lanes $3 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.readlines()[-1])
e = d['e2e']
print('  e2e %.1f k images/s  %.4f ms/step  %.1f GB/s  equal=%s' % (e['value'] / 1e3, e['ms_per_step'], e['h2d_bytes_per_step'] / e['ms_per_step'] / 1e6, e.get('outputs_equal_to_resident_run')))
"
done
