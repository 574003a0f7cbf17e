#!/bin/bash
# Round 2, first GPU call: the new streaming ring kernel on hardware for the first time.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "[$((SECONDS-T0))s] $*" | tee -a gpurun_out/c1.log; }
note "start $(nvidia-smi --query-gpu=name,driver_version --format=csv,noheader | head -1)"
note "toolchains: cargo=$(which cargo 2>/dev/null) rustc=$(which rustc 2>/dev/null) nproc=$(nproc) cpu.max=$(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1_smoke.log 2>&1; note "smoke rc=$? $(tail -2 gpurun_out/c1_smoke.log | tr '\n' ' ' | head -c 600)"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/c1_parity.log 2>&1; note "parity rc=$? $(tail -3 gpurun_out/c1_parity.log | tr '\n' ' ' | head -c 900)"
timeout 300 python bench.py --steps 6 --no-cpu --no-e2e > gpurun_out/c1_bench_c2.json 2>gpurun_out/c1_bench_c2.err; note "bench c2 rc=$? $(head -c 1500 gpurun_out/c1_bench_c2.json)"
timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --min-items 8192 > gpurun_out/c1_bench_c2_nb2.json 2>&1; note "bench c2 items8192 $(python tools/kms.py gpurun_out/c1_bench_c2_nb2.json)"
timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --min-items 4096 > gpurun_out/c1_bench_c2_nb1.json 2>&1; note "bench c2 items4096 $(python tools/kms.py gpurun_out/c1_bench_c2_nb1.json)"
timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --min-items 16384 > gpurun_out/c1_bench_c2_nb4.json 2>&1; note "bench c2 items16384 $(python tools/kms.py gpurun_out/c1_bench_c2_nb4.json)"
timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --strip-cols 64 > gpurun_out/c1_bench_c2_s64.json 2>&1; note "bench c2 strip64 $(python tools/kms.py gpurun_out/c1_bench_c2_s64.json)"
timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --workload c2_4k_to_512_lanczos3 > gpurun_out/c1_bench_l3.json 2>&1; note "bench lanczos3 $(python tools/kms.py gpurun_out/c1_bench_l3.json)"
timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --alpha 1 > gpurun_out/c1_bench_alpha.json 2>&1; note "bench c2 alpha $(python tools/kms.py gpurun_out/c1_bench_alpha.json)"
timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --workload c3_8k_to_1080p_robidoux_sharpen > gpurun_out/c1_bench_c3.json 2>&1; note "bench c3 $(python tools/kms.py gpurun_out/c1_bench_c3.json)"
timeout 200 python bench.py --steps 4 --no-cpu --no-e2e --workload c4_1080p_to_4k_mitchell_sepia_over --batch 128 > gpurun_out/c1_bench_c4.json 2>&1; note "bench c4 $(python tools/kms.py gpurun_out/c1_bench_c4.json)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hv_ring -s 2 -c 1 -o gpurun_out/prof_hv_c1 python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu --no-e2e --no-check > gpurun_out/c1_ncu.log 2>&1; note "ncu rc=$? $(ls -la gpurun_out/prof_hv_c1.ncu-rep 2>&1 | head -c 200)"
timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity.py > gpurun_out/c1_rest.log 2>&1; note "rest rc=$? $(tail -3 gpurun_out/c1_rest.log | tr '\n' ' ' | head -c 600)"
note "end"
