#!/bin/bash
# gpurun driver: timing-fuzzer campaign (clock-spin delays at every synchronisation site of the tcgen05 kernels)
cd "${GRAFT_REPO_ROOT:-.}"
L=kws_b200/lib/libfastgrnn_b200.so
F=kws_b200/lib/libfastgrnn_b200_fuzz.so
P=tools/first_launch_probe
O=gpurun_out/hunt5
mkdir -p $O; rm -f $O/*
( timeout 900 $P $F loop 64 400 1 1 ) > $O/fuzz64.log 2>&1
( timeout 900 $P $F loop 2048 150 1 1 ) > $O/fuzz2048.log 2>&1
( timeout 900 $P $F loop 8192 60 0 1 ) > $O/fuzz8192.log 2>&1
( FGRNN_TC_NT=4 timeout 900 $P $F loop 8192 40 0 1 ) > $O/fuzz8192_nt4.log 2>&1
( FGRNN_TC_NS=32 timeout 900 $P $F loop 200 200 1 1 ) > $O/fuzz200_ns32.log 2>&1
for i in $(seq 1 20); do PROBE_SEED=$i timeout 120 $P $F fresh 64 0 1 1; done > $O/fuzz_fresh64.log 2>&1
tail -q -n 1 $O/fuzz*.log
grep -c "fresh ok" $O/fuzz_fresh64.log
# the GPU suite on the fuzzed library (forward, BPTT recurrence and contraction kernels all carry the fuzzer)
KWS_B200_LIB=$PWD/$F timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_fuzz.log 2>&1; echo "pytest(fuzz) rc=$?"; tail -n 4 $O/pytest_fuzz.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_quiet.log 2>&1; echo "pytest(quiet) rc=$?"; tail -n 4 $O/pytest_quiet.log
