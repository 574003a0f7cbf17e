#!/bin/bash
# Round 2, fourth GPU call: two row-streams per lane (band pairs).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "[$((SECONDS-T0))s] $*" | tee -a gpurun_out/c4.log; }
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/c4_parity.log 2>&1; note "parity rc=$? $(tail -3 gpurun_out/c4_parity.log | tr '\n' ' ' | head -c 900)"
timeout 300 python bench.py --steps 6 --no-cpu --no-e2e --no-others > gpurun_out/c4_bench_c2.json 2>gpurun_out/c4_bench_c2.err; note "bench c2 $(python tools/kms.py gpurun_out/c4_bench_c2.json)"
for tag in pf0; do IFB200_LIB=$PWD/imageflow_b200/libifb200_$tag.so timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --no-others > gpurun_out/c4_bench_$tag.json 2>&1; note "bench c2 $tag $(python tools/kms.py gpurun_out/c4_bench_$tag.json)"; done
for mi in 8192 16384 32768; do timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --no-others --min-items $mi > gpurun_out/c4_bench_mi$mi.json 2>&1; note "bench c2 items$mi $(python tools/kms.py gpurun_out/c4_bench_mi$mi.json)"; done
timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --no-others --workload c2_4k_to_512_lanczos3 > gpurun_out/c4_bench_l3.json 2>&1; note "bench lanczos3 $(python tools/kms.py gpurun_out/c4_bench_l3.json)"
timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --no-others --alpha 1 > gpurun_out/c4_bench_alpha.json 2>&1; note "bench c2 alpha $(python tools/kms.py gpurun_out/c4_bench_alpha.json)"
timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --no-others --workload c3_8k_to_1080p_robidoux_sharpen > gpurun_out/c4_bench_c3.json 2>&1; note "bench c3 $(python tools/kms.py gpurun_out/c4_bench_c3.json)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hv_ring -s 2 -c 1 -o gpurun_out/prof_hv_c4 python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu --no-e2e --no-check --no-others > gpurun_out/c4_ncu.log 2>&1; note "ncu rc=$? $(ls -la gpurun_out/prof_hv_c4.ncu-rep 2>&1 | head -c 200)"
timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity.py > gpurun_out/c4_rest.log 2>&1; note "rest rc=$? $(tail -3 gpurun_out/c4_rest.log | tr '\n' ' ' | head -c 600)"
note "end"
