#!/bin/bash
# usage: tools/quickbench.sh <tag> [bench args...]   -> prints one summary line
tag=$1; shift
timeout 300 python bench.py --batch 256 --steps 8 --no-cpu --no-e2e "$@" 2>&1 | tail -1 > gpurun_out/qb_$tag.log
python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/qb_{tag}.log").read())
    print(f"{tag:28s} {d['ms_per_step']:8.3f} ms  frac {d['roofline']['frac']:.3f}  {d['value']/1e3:8.1f} Gpx/s  check {d['parity_check']}  fused {d['roofline']['fused_jobs']} clk {d['clocks']['sm_mhz']} pw {d['clocks'].get('power_w_max')}")
except Exception as e:
    print(tag, "FAILED", e, open(f"gpurun_out/qb_{tag}.log").read()[-300:])
PY
