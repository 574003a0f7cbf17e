#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
echo "== plain" > gpurun_out/dbg2.log; timeout 300 python tools/r2_dbg2.py >> gpurun_out/dbg2.log 2>&1
for tool in memcheck initcheck synccheck; do echo "== $tool" >> gpurun_out/dbg2.log; timeout 600 compute-sanitizer --tool $tool --print-limit 10 python tools/r2_dbg2.py >> gpurun_out/dbg2.log 2>&1; done
tail -60 gpurun_out/dbg2.log
