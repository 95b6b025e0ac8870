#!/bin/bash
# Developer tool: A/B the step time of variants in ONE gpurun call (box-to-box noise is ~1 us, in-call noise ~0.1 us).
#   tools/ab_variants.sh "<nvcc flags> :: <bench args>" ...   -> one line per variant and round
mkdir -p gpurun_out
for r in 1 2; do
  i=0
  for v in "$@"; do
    export SIHL_B200_NVCC_EXTRA="${v%%::*}"
    args="${v#*::}"; [ "$args" = "$v" ] && args=""
    python bench.py --skip-cpu-baseline --skip-gpu-eager --skip-e2e $args > gpurun_out/ab_${i}_${r}.json 2> gpurun_out/ab_${i}_${r}.err
    i=$((i+1))
  done
done
python - "$@" <<'PY'
import json, sys
for i, f in enumerate(sys.argv[1:]):
    for r in (1, 2):
        try:
            d = json.load(open(f"gpurun_out/ab_{i}_{r}.json"))
            o = d.get("other_decode_mode") or {}
            print(f"[{f or 'default'}] run {r}: {d['execution']['decode_mode']} {d['ms_per_step']*1e3:.2f} us  other {o.get('ms_per_step', 0)*1e3:.2f} us")
        except Exception as e:
            print(f"[{f}] run {r}: failed {e}")
PY
