#!/bin/bash
# First GPU call of the next round (everything that was written after round 1's GPU budget was spent):
#   1. the whole GPU suite -- tests/test_zz_whitespace_gpu.py should report XPASS (then drop its xfail marker);
#   2. config 5 with the enqueue profile (ifb200_batch_host_profile): where the 127 ms of host time go;
#   3. whitespace code-map kernel: time per 4K frame (CUDA events) -- expect an HBM-bound 5 bytes per pixel;
#   4. the bench line.
# One .ncu-rep per call at most (gpurun_out is capped at 64 MiB).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "[$((SECONDS-T0))s] $*" | tee -a gpurun_out/r2_first.log; }
note "start"
timeout 200 python -m pytest tests -q -m gpu -rxX > gpurun_out/r2_pytest.log 2>&1; note "pytest rc=$? $(tail -1 gpurun_out/r2_pytest.log)"; grep -i "xpass\|xfail" gpurun_out/r2_pytest.log | tee -a gpurun_out/r2_first.log
timeout 150 python tools/mixed_workload.py --images 2000 --check 0 > gpurun_out/r2_c5.json 2>gpurun_out/r2_c5.err; note "c5 rc=$? $(tail -1 gpurun_out/r2_c5.json | head -c 1200)"
timeout 100 python - <<'PY' 2>&1 | tee -a gpurun_out/r2_first.log
import torch, numpy as np
import imageflow_b200 as ifb
from imageflow_b200 import synth
b = ifb.Batch(0)
imgs = [synth.noise_torch(3840, 2160, seed=i, alpha_mode="mixed") for i in range(64)]       # 2.1 GB: far larger than L2
st = torch.cuda.current_stream().cuda_stream
for t in imgs[:4]:
    b.detect_content(ifb.BitmapWindow.from_torch(t, alpha_meaningful=True), 1, stream=st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
t0 = time.perf_counter(); e0.record()
for t in imgs:
    r = b.detect_content(ifb.BitmapWindow.from_torch(t, alpha_meaningful=True), 1, stream=st)
e1.record(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"detect_content, 64 x 4K noise frames: {dt/64*1e3:.3f} ms per frame wall (kernel + 8.3 MB D2H + host walk), last rect {r}")
PY
timeout 240 python bench.py > gpurun_out/r2_bench.json 2>gpurun_out/r2_bench.err; note "bench rc=$? $(head -c 400 gpurun_out/r2_bench.json)"
note "end"
