#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/misc1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py > gpurun_out/misc1/dist_check.txt 2>&1; echo "dist rc=$?"; grep -E "DIST_CHECK|Error|assert" gpurun_out/misc1/dist_check.txt | head -5
timeout 300 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2
timeout 120 python tools/peaks_probe.py > gpurun_out/misc1/peaks.json 2> gpurun_out/misc1/peaks.err; cat gpurun_out/misc1/peaks.json
