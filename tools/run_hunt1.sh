#!/bin/bash
# gpurun driver: first-launch / stale-state hunt for the tcgen05 forward (round 2)
cd "${GRAFT_REPO_ROOT:-.}"
L=kws_b200/lib/libfastgrnn_b200.so
F=kws_b200/lib/libfastgrnn_b200_fuzz.so
P=tools/first_launch_probe
O=gpurun_out/hunt1
mkdir -p $O
nvidia-smi -q | grep -i -m2 "persistence" > $O/env.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $O/env.log 2>&1
# 1. fresh processes first (the box is at its freshest now)
for i in $(seq 1 60); do PROBE_SEED=$i timeout 120 $P $L fresh 64 0 1 1; done > $O/fresh64_poison.log 2>&1
for i in $(seq 1 25); do PROBE_SEED=$i timeout 120 $P $L fresh 8192 0 0 1; done > $O/fresh8192_poison.log 2>&1
for i in $(seq 1 40); do PROBE_SEED=$i timeout 120 $P $L fresh 64 0 1 0; done > $O/fresh64_plain.log 2>&1
for i in $(seq 1 15); do PROBE_SEED=$i timeout 120 $P $L fresh 8192 0 0 0; done > $O/fresh8192_plain.log 2>&1
# 2. in-process alternating weights with poison
timeout 600 $P $L loop 64 1500 1 1 > $O/loop64.log 2>&1
timeout 600 $P $L loop 2048 400 1 1 > $O/loop2048.log 2>&1
timeout 600 $P $L loop 8192 150 0 1 > $O/loop8192.log 2>&1
FGRNN_TC_NT=4 timeout 600 $P $L loop 8192 80 0 1 > $O/loop8192_nt4.log 2>&1
# 3. timing fuzzer build
timeout 600 $P $F loop 64 300 1 1 > $O/fuzz64.log 2>&1
timeout 600 $P $F loop 2048 100 1 1 > $O/fuzz2048.log 2>&1
timeout 600 $P $F loop 8192 40 0 1 > $O/fuzz8192.log 2>&1
FGRNN_TC_NT=4 timeout 600 $P $F loop 8192 30 0 1 > $O/fuzz8192_nt4.log 2>&1
tail -n 3 $O/*.log
# 4. the GPU suite as it stands
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 $O/pytest.log
