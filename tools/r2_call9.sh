#!/bin/bash
# Round 2, GPU call: A/B of builds of the ring kernel (libifb200_<tag>.so), parity of the default one, ncu of the default one.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T0=$SECONDS
TAG=${1:-c9}; shift
note() { echo "[$((SECONDS-T0))s] $*" | tee -a gpurun_out/$TAG.log; }
B="python bench.py --no-cpu --no-e2e --no-others"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/${TAG}_parity.log 2>&1; note "parity rc=$? $(tail -3 gpurun_out/${TAG}_parity.log | tr '\n' ' ' | head -c 900)"
for round in 1 2; do
timeout 300 $B --steps 6 > gpurun_out/${TAG}_bench_c2.json 2>gpurun_out/${TAG}_bench_c2.err; note "bench c2 default $(python tools/kms.py gpurun_out/${TAG}_bench_c2.json)"
for t in "$@"; do IFB200_LIB=$PWD/imageflow_b200/libifb200_$t.so timeout 300 $B --steps 6 > gpurun_out/${TAG}_bench_c2_$t.json 2>&1; note "bench c2 $t $(python tools/kms.py gpurun_out/${TAG}_bench_c2_$t.json)"; done
done
for mi in 16384 24576; do timeout 200 $B --steps 4 --min-items $mi > gpurun_out/${TAG}_mi$mi.json 2>&1; note "bench c2 items$mi $(python tools/kms.py gpurun_out/${TAG}_mi$mi.json)"; done
timeout 200 $B --steps 4 --workload c2_4k_to_512_lanczos3 > gpurun_out/${TAG}_bench_l3.json 2>&1; note "bench lanczos3 $(python tools/kms.py gpurun_out/${TAG}_bench_l3.json)"
timeout 200 $B --steps 4 --alpha 1 > gpurun_out/${TAG}_bench_alpha.json 2>&1; note "bench c2 alpha $(python tools/kms.py gpurun_out/${TAG}_bench_alpha.json)"
timeout 200 $B --steps 4 --workload c3_8k_to_1080p_robidoux_sharpen > gpurun_out/${TAG}_bench_c3.json 2>&1; note "bench c3 $(python tools/kms.py gpurun_out/${TAG}_bench_c3.json)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hv_ring -s 2 -c 1 -o gpurun_out/prof_hv_$TAG $B --batch 256 --steps 1 --warmup 1 --no-check > gpurun_out/${TAG}_ncu.log 2>&1; note "ncu rc=$? $(ls -la gpurun_out/prof_hv_$TAG.ncu-rep 2>&1 | head -c 200)"
note "end"
