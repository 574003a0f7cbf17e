#!/bin/bash
# compute-sanitizer initcheck + synccheck on the ring-kernel and tile-kernel parity tests (racecheck + memcheck are in r2_final.sh)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for tool in initcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_zz_whitespace_gpu.py -q -x -k "ring_kernel_forms or tile_kernel_bit_exact or fused_decompositions or detect_content" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "compute-sanitizer $tool rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitizer_$tool.log | tail -3 | tr '\n' ' ' | head -c 400)"
done
