#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e --no-others --no-check"
for round in 1 2; do
for t in "" "$@"; do
  if [ -z "$t" ]; then L=""; n=default; else L="IFB200_LIB=$PWD/imageflow_b200/libifb200_$t.so"; n=$t; fi
  env $L timeout 300 $B --steps 6 --workload c4_1080p_to_4k_mitchell_sepia_over --batch 128 > gpurun_out/ab_c4_$n.json 2>gpurun_out/ab_c4_$n.err; echo "c4/128 $n $(python tools/kms.py gpurun_out/ab_c4_$n.json)"
done
done
