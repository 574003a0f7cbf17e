#!/bin/bash
# config 5 on one GPU: parity of the test suite's ring cases, the 2 000-image run (twice), per-launch kernel times under ncu, the bench line's kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/c18_gpu.log 2>&1; tail -1 gpurun_out/c18_gpu.log
for k in 1 2; do timeout 600 python tools/mixed_workload.py --images 2000 --check 4 > gpurun_out/c5_2000_1gpu.json 2> gpurun_out/c5_2000_1gpu.err; tail -c 900 gpurun_out/c5_2000_1gpu.json; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hv_ring -c 1200 --csv --log-file gpurun_out/c5_launches.csv python tools/mixed_workload.py --images 400 --check 0 > gpurun_out/c5_ncu.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/c5_launches.csv")) if len(r)>5]
h=rows[0]; ni=h.index("Kernel Name"); vi=h.index("Metric Value"); gi=h.index("Grid Size")
t=[(float(r[vi].replace(",","")), r[gi]) for r in rows[1:] if "hv_ring" in r[ni]]
d=sorted(x[0] for x in t)
print(len(t), "hv launches; sum ms", sum(d)/1e6, "mean us", sum(d)/max(len(d),1)/1e3, "min/median/p90/max us", d[0]/1e3, d[len(d)//2]/1e3, d[int(len(d)*0.9)]/1e3, d[-1]/1e3)
PY
python bench.py --no-cpu --no-e2e --no-others --steps 5 | python tools/kms.py /dev/stdin
