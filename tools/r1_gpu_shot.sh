#!/bin/bash
# One gpurun call at the end of round 1 (GPU budget: ~11 minutes).  Most important first; every step has its own timeout.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "[$((SECONDS-T0))s] $*" | tee -a gpurun_out/shot.log; }
note "start"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | tee -a gpurun_out/shot.log
# 1. new code first: both tile-kernel forms, extreme down-scales
timeout 200 python -m pytest tests -q -m gpu -k "tile_kernel or extreme" --durations=5 > gpurun_out/pytest_new.log 2>&1; note "pytest new rc=$? $(tail -1 gpurun_out/pytest_new.log)"
# 2. everything else
timeout 330 python -m pytest tests -q -m gpu -k "not (tile_kernel or extreme)" --durations=8 > gpurun_out/pytest_rest.log 2>&1; note "pytest rest rc=$? $(tail -1 gpurun_out/pytest_rest.log)"
# 3. config 4: first form vs second form of the tile kernel
for f in 1 2; do
  timeout 120 python bench.py --workload c4_1080p_to_4k_mitchell_sepia_over --batch 128 --steps 6 --no-cpu --no-e2e --tile-kernel $f 2>gpurun_out/c4_tile$f.err | tail -1 > gpurun_out/c4_tile$f.json
  note "c4 tile$f: $(python -c "import json;d=json.load(open('gpurun_out/c4_tile$f.json'));print(d['roofline']['kernel_ms'], d['roofline']['kernel_ms_min'], round(d['roofline']['frac'],4), d['parity_check'], d['clocks']['sm_mhz'])" 2>&1 | tail -1)"
done
# 4. the bench line (config 2, defaults), as the driver runs it
timeout 240 python bench.py > gpurun_out/bench_r1_final.json 2>gpurun_out/bench_r1_final.err; note "bench rc=$? $(head -c 600 gpurun_out/bench_r1_final.json)"
# 5. ncu: full capture of the second tile kernel (config 4, 8 frames), then the launch list of the same command
timeout 150 ncu --set full --clock-control none --import-source on -k regex:fused_tile2 -c 1 -f -o gpurun_out/r1_tile2 \
  python bench.py --workload c4_1080p_to_4k_mitchell_sepia_over --batch 8 --steps 1 --warmup 1 --no-cpu --no-e2e --no-check > gpurun_out/ncu_tile2.log 2>&1; note "ncu tile2 rc=$?"
# 6. config 5 (mixed thumbnails) after the host-side plan work
timeout 150 python tools/mixed_workload.py --images 2000 --check 0 > gpurun_out/c5_2000.json 2>gpurun_out/c5_2000.err; note "c5 rc=$? $(tail -1 gpurun_out/c5_2000.json | head -c 700)"
note "end"
