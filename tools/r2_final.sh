#!/bin/bash
# Round 2, closing GPU calls: whole GPU test suite, smoke(), the default bench line (both arms), DRAM traffic of the shipped kernel,
# launch list, compute-sanitizer on the ring-kernel and tile-kernel parity tests.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T0=$SECONDS
TAG=${1:-fin}
note() { echo "[$((SECONDS-T0))s] $*" | tee -a gpurun_out/$TAG.log; }
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_gpu_tests.log 2>&1; note "gpu tests rc=$? $(tail -3 gpurun_out/${TAG}_gpu_tests.log | tr '\n' ' ' | head -c 600)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; note "smoke rc=$? $(tail -1 gpurun_out/${TAG}_smoke.log | head -c 300)"
timeout 900 python bench.py > gpurun_out/${TAG}_bench_line.json 2>gpurun_out/${TAG}_bench_line.err; note "bench rc=$? $(python tools/kms.py gpurun_out/${TAG}_bench_line.json)"
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2>gpurun_out/${TAG}_bench_ref.err; note "bench reference rc=$? $(tail -c 400 gpurun_out/${TAG}_bench_ref.json)"
timeout 600 python bench.py --ncu-traffic --no-cpu --no-e2e --no-others --steps 2 > gpurun_out/${TAG}_traffic.log 2>&1; note "traffic rc=$? $(cat profiles/traffic_c2_4k_to_512_robidoux.json | tr '\n' ' ' | head -c 400)"
for w in c4_1080p_to_4k_mitchell_sepia_over c2_4k_to_512_lanczos3 c3_8k_to_1080p_robidoux_sharpen; do
  timeout 300 python bench.py --ncu-traffic --workload $w --no-cpu --no-e2e --no-others --steps 2 > gpurun_out/${TAG}_traffic_$w.log 2>&1; note "traffic $w rc=$? $(cat profiles/traffic_$w.json | tr '\n' ' ' | head -c 300)"
done
cp profiles/traffic_*.json gpurun_out/ 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-others --no-check > gpurun_out/${TAG}_launches.log 2>&1; note "launch list rc=$? $(wc -l < gpurun_out/${TAG}_launches.csv) lines"
for tool in racecheck memcheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "ring_kernel_forms or tile_kernel_bit_exact or fused_decompositions" > gpurun_out/${TAG}_sanitizer_$tool.log 2>&1; note "compute-sanitizer $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/${TAG}_sanitizer_$tool.log | tail -3 | tr '\n' ' ' | head -c 400)"
done
note "end"
