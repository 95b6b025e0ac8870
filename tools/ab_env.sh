#!/bin/bash
# Developer tool: A/B runtime knobs (env vars / bench args) in ONE gpurun call.  tools/ab_env.sh "ENV=.. :: args" ...
mkdir -p gpurun_out
for r in 1 2; do i=0; for v in "$@"; do
  envs="${v%%::*}"; args="${v#*::}"; [ "$args" = "$v" ] && args=""
  env $envs python bench.py --skip-cpu-baseline --skip-gpu-eager --skip-e2e --steps 1500 --warmup 100 $args > gpurun_out/abe_${i}_$r.json 2>/dev/null; i=$((i+1)); done; done
python - "$@" <<'PY'
import json, sys
for i, f in enumerate(sys.argv[1:]):
    for r in (1, 2):
        try:
            d = json.load(open(f"gpurun_out/abe_{i}_{r}.json")); o = d.get("other_decode_mode") or {}
            print(f"[{f}] run {r}: dense {d['ms_per_step']*1e3:.2f} us  decode alone {d['roofline']['kernel_ms']*1e3:.2f} us  candidate-first {o.get('ms_per_step',0)*1e3:.2f} us")
        except Exception as e: print(f"[{f}] run {r}: failed {e}")
PY
