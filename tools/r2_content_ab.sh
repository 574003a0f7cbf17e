#!/bin/bash
# config 4 (tile kernel) on noise and on gradient content: the table gathers of the epilogue conflict less on smooth content
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e --no-others"
for c in noise gradient; do
  timeout 300 $B --steps 6 --workload c4_1080p_to_4k_mitchell_sepia_over --batch 128 --content $c > gpurun_out/c4_$c.json 2>gpurun_out/c4_$c.err; echo "c4/128 $c $(python tools/kms.py gpurun_out/c4_$c.json)"
  timeout 300 $B --steps 6 --content $c > gpurun_out/c2_$c.json 2>gpurun_out/c2_$c.err; echo "c2/1024 $c $(python tools/kms.py gpurun_out/c2_$c.json)"
done
