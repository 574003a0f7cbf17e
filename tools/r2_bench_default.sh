#!/bin/bash
# the default bench line (what the driver runs), with its wall-clock time
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
T0=$SECONDS
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$? wall $((SECONDS-T0)) s"
python tools/kms.py gpurun_out/bench_default.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
for k, v in (d.get("other_workloads") or {}).items():
    print(k, v.get("images_per_step"), v.get("roofline", {}).get("frac"), v.get("roofline", {}).get("kernel_ms"), v.get("parity_check"), v.get("error"))
print("traffic", d["roofline"].get("traffic"), d["roofline"].get("traffic_source"))
PY
tail -3 gpurun_out/bench_default.err
