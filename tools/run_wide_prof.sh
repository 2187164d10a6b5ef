#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/wide
python tools/prof_wide.py 64 256 8192 > gpurun_out/wide/plain1.log 2>&1 && python tools/prof_wide.py 256 128 8192 > gpurun_out/wide/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_ -s 2 -c 2 -o gpurun_out/wide/prof_l1 -f python tools/prof_wide.py 64 256 8192 > gpurun_out/wide/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_ -s 2 -c 2 -o gpurun_out/wide/prof_l2 -f python tools/prof_wide.py 256 128 8192 > gpurun_out/wide/ncu2.log 2>&1
ls -la gpurun_out/wide; tail -3 gpurun_out/wide/ncu1.log
