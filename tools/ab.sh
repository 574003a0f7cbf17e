#!/bin/bash
# usage: ab.sh tag1 tag2 ...  (libs imageflow_b200/libifb200_<tag>.so), prints kernel_ms_min per run, two rounds
for round in 1 2; do for tag in "$@"; do
IFB200_LIB=$PWD/imageflow_b200/libifb200_$tag.so timeout 300 python bench.py --batch 1024 --steps 6 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/ab_$tag.log
python - "$tag" <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/ab_{t}.log").read()); r=d['roofline']
    print(f"{t:10s} min {r['kernel_ms_min']:.3f} mean {r['kernel_ms']:.3f} clk {d['clocks']['sm_mhz']} pw {d['clocks'].get('power_w_max')} chk {d['parity_check']['max_abs_delta_vs_oracle']}")
except Exception as e: print(t,'FAILED',e)
PY
done; done
