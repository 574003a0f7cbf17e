#!/bin/bash
# second end-of-round GPU call: tile kernel (second form, slot H pass) and the ring kernel's gather-ahead form
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "[$((SECONDS-T0))s] $*" | tee -a gpurun_out/shot2.log; }
note "start"
timeout 200 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/pytest2.log 2>&1; note "pytest rc=$? $(tail -1 gpurun_out/pytest2.log)"
C4="--workload c4_1080p_to_4k_mitchell_sepia_over --batch 128 --steps 6 --no-cpu --no-e2e"
for f in 1 2; do
  timeout 120 python bench.py $C4 --tile-kernel $f 2>gpurun_out/c4b_tile$f.err | tail -1 > gpurun_out/c4b_tile$f.json
  note "c4 tile$f: $(python -c "import json;d=json.load(open('gpurun_out/c4b_tile$f.json'));print(d['roofline']['kernel_ms'], d['roofline']['kernel_ms_min'], round(d['roofline']['frac'],4), d['parity_check'], d['clocks']['sm_mhz'])" 2>&1 | tail -1)"
done
for round in 1 2; do for ga in 0 1; do
  timeout 120 python bench.py --batch 1024 --steps 8 --no-cpu --no-e2e --gather-ahead $ga 2>gpurun_out/c2_ga$ga.err | tail -1 > gpurun_out/c2_ga${ga}_r$round.json
  note "c2 ga=$ga: $(python -c "import json;d=json.load(open('gpurun_out/c2_ga${ga}_r$round.json'));print(d['roofline']['kernel_ms'], d['roofline']['kernel_ms_min'], round(d['roofline']['frac'],4), d['parity_check'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))" 2>&1 | tail -1)"
done; done
for ga in 0 1; do
  timeout 120 python bench.py --batch 512 --steps 8 --no-cpu --no-e2e --alpha 1 --gather-ahead $ga 2>>gpurun_out/c2_ga$ga.err | tail -1 > gpurun_out/c2a_ga$ga.json
  note "c2 alpha ga=$ga: $(python -c "import json;d=json.load(open('gpurun_out/c2a_ga$ga.json'));print(d['roofline']['kernel_ms'], d['roofline']['kernel_ms_min'], round(d['roofline']['frac'],4), d['parity_check'])" 2>&1 | tail -1)"
done
timeout 120 ncu --set full --clock-control none --import-source on -k regex:fused_tile2 -c 1 -f -o gpurun_out/r1_tile2_v2 \
  python bench.py $C4 --batch 8 --steps 1 --warmup 1 --no-check > gpurun_out/ncu_tile2_v2.log 2>&1; note "ncu tile2 rc=$?"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:fused_down -c 1 -f -o gpurun_out/r1_fused_ga \
  python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu --no-e2e --no-check --gather-ahead 1 > gpurun_out/ncu_ga.log 2>&1; note "ncu ga rc=$?"
note "end"
