#!/bin/bash
# BASELINE.json configs[4] at its stated scale: 10 000 images, export_4_sizes cascades, sharded over the 8 GPUs of one box.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/mixed_workload.py --images 10000 --check 8 > gpurun_out/c5_${N}gpu.json 2> gpurun_out/c5_${N}gpu.err
echo "rc=$? $(tail -c 1500 gpurun_out/c5_${N}gpu.json)"; tail -5 gpurun_out/c5_${N}gpu.err
timeout 600 python tools/mixed_workload.py --images 2000 --check 4 > gpurun_out/c5_2000_1gpu.json 2> gpurun_out/c5_2000_1gpu.err
echo "1 GPU 2000 images rc=$? $(tail -c 1200 gpurun_out/c5_2000_1gpu.json)"
