#!/bin/bash
# Round 2, second GPU call: which launches of the ring kernel trap, with and without 16-byte aligned TMA box origins.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "[$((SECONDS-T0))s] $*" | tee -a gpurun_out/c2.log; }
note "== aligned strip origins (default)"
timeout 600 python tools/r2_dbg.py 2>&1 | tee -a gpurun_out/c2.log
note "== unaligned strip origins (IFB200_DEBUG_K0_ALIGN=1)"
IFB200_DEBUG_K0_ALIGN=1 timeout 600 python tools/r2_dbg.py 2>&1 | tee -a gpurun_out/c2.log
note "== sanitizer on case D, unaligned"
IFB200_DEBUG_K0_ALIGN=1 timeout 300 compute-sanitizer --tool memcheck python tools/r2_dbg.py one '{"name":"D","iw":640,"ih":480,"ow":200,"oh":150,"alpha":0}' > gpurun_out/c2_sanitizer.log 2>&1; note "sanitizer rc=$? $(grep -m3 -A8 '=========' gpurun_out/c2_sanitizer.log | head -40 | tr '\n' '|' | head -c 1500)"
note "== full parity suite + bench with the default"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/c2_parity.log 2>&1; note "parity rc=$? $(tail -3 gpurun_out/c2_parity.log | tr '\n' ' ' | head -c 900)"
timeout 300 python bench.py --steps 6 --no-cpu --no-e2e > gpurun_out/c2_bench_c2.json 2>gpurun_out/c2_bench_c2.err; note "bench c2 $(python tools/kms.py gpurun_out/c2_bench_c2.json)"
for mi in 4096 8192 16384; do timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --min-items $mi > gpurun_out/c2_bench_mi$mi.json 2>&1; note "bench c2 items$mi $(python tools/kms.py gpurun_out/c2_bench_mi$mi.json)"; done
timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --strip-cols 64 > gpurun_out/c2_bench_s64.json 2>&1; note "bench c2 strip64 $(python tools/kms.py gpurun_out/c2_bench_s64.json)"
timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --workload c2_4k_to_512_lanczos3 > gpurun_out/c2_bench_l3.json 2>&1; note "bench lanczos3 $(python tools/kms.py gpurun_out/c2_bench_l3.json)"
timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --alpha 1 > gpurun_out/c2_bench_alpha.json 2>&1; note "bench c2 alpha $(python tools/kms.py gpurun_out/c2_bench_alpha.json)"
timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --workload c3_8k_to_1080p_robidoux_sharpen > gpurun_out/c2_bench_c3.json 2>&1; note "bench c3 $(python tools/kms.py gpurun_out/c2_bench_c3.json)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hv_ring -s 2 -c 1 -o gpurun_out/prof_hv_c2 python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu --no-e2e --no-check > gpurun_out/c2_ncu.log 2>&1; note "ncu rc=$? $(ls -la gpurun_out/prof_hv_c2.ncu-rep 2>&1 | head -c 200)"
timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity.py > gpurun_out/c2_rest.log 2>&1; note "rest rc=$? $(tail -3 gpurun_out/c2_rest.log | tr '\n' ' ' | head -c 600)"
note "end"
