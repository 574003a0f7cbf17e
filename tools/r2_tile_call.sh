#!/bin/bash
# Round 2, GPU call for the tile kernel (config 4) and the section 8(f) neighbours: parity, bench, ncu of fused_tile2_kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T0=$SECONDS
TAG=${1:-t1}; shift
note() { echo "[$((SECONDS-T0))s] $*" | tee -a gpurun_out/$TAG.log; }
B="python bench.py --no-cpu --no-e2e --no-others"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_whitespace_gpu.py -x -q -m gpu > gpurun_out/${TAG}_parity.log 2>&1; note "parity rc=$? $(tail -3 gpurun_out/${TAG}_parity.log | tr '\n' ' ' | head -c 900)"
for round in 1 2; do
timeout 300 $B --steps 6 --workload c4_1080p_to_4k_mitchell_sepia_over --batch 128 > gpurun_out/${TAG}_bench_c4.json 2>gpurun_out/${TAG}_bench_c4.err; note "bench c4/128 default $(python tools/kms.py gpurun_out/${TAG}_bench_c4.json)"
for t in "$@"; do IFB200_LIB=$PWD/imageflow_b200/libifb200_$t.so timeout 300 $B --steps 6 --workload c4_1080p_to_4k_mitchell_sepia_over --batch 128 > gpurun_out/${TAG}_bench_c4_$t.json 2>&1; note "bench c4/128 $t $(python tools/kms.py gpurun_out/${TAG}_bench_c4_$t.json)"; done
done
timeout 300 $B --steps 4 --workload c4_1080p_to_4k_mitchell_sepia_over > gpurun_out/${TAG}_bench_c4_full.json 2>&1; note "bench c4/512 $(python tools/kms.py gpurun_out/${TAG}_bench_c4_full.json)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_tile2 -s 2 -c 1 -o gpurun_out/prof_tile2_$TAG $B --workload c4_1080p_to_4k_mitchell_sepia_over --batch 16 --steps 1 --warmup 1 --no-check > gpurun_out/${TAG}_ncu.log 2>&1; note "ncu rc=$? $(ls -la gpurun_out/prof_tile2_$TAG.ncu-rep 2>&1 | head -c 200)"
if [ -z "$SKIP_NB" ]; then timeout 600 python tools/bench_neighbours.py > gpurun_out/neighbours_r2.jsonl 2> gpurun_out/neighbours_r2.err; note "neighbours rc=$? $(wc -l < gpurun_out/neighbours_r2.jsonl) lines $(tail -c 300 gpurun_out/neighbours_r2.err)"; fi
note "end"
