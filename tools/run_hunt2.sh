#!/bin/bash
# gpurun driver: timing-fuzzer build of the tcgen05 kernels + compute-sanitizer racecheck on the smallest case
cd "${GRAFT_REPO_ROOT:-.}"
L=kws_b200/lib/libfastgrnn_b200.so
F=kws_b200/lib/libfastgrnn_b200_fuzz.so
P=tools/first_launch_probe
O=gpurun_out/hunt2
mkdir -p $O
( time timeout 900 $P $F loop 64 400 1 1 ) > $O/fuzz64.log 2>&1
( time timeout 900 $P $F loop 2048 120 1 1 ) > $O/fuzz2048.log 2>&1
( time timeout 900 $P $F loop 8192 40 0 1 ) > $O/fuzz8192.log 2>&1
( FGRNN_TC_NT=4 timeout 900 $P $F loop 8192 30 0 1 ) > $O/fuzz8192_nt4.log 2>&1
( FGRNN_TC_NS=32 timeout 900 $P $F loop 200 200 1 1 ) > $O/fuzz200_ns32.log 2>&1
tail -n 4 $O/fuzz*.log
# fuzzed training step through python (forward + BPTT kernels), compared with the quiet library
KWS_B200_LIB=$PWD/$F timeout 900 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_first_launch.py -x -q > $O/pytest_fuzz.log 2>&1; echo "pytest(fuzz) rc=$?"; tail -n 5 $O/pytest_fuzz.log
timeout 900 python -m pytest tests/test_gpu_first_launch.py tests/test_gpu_tcgen05.py -x -q > $O/pytest_quiet.log 2>&1; echo "pytest(quiet) rc=$?"; tail -n 5 $O/pytest_quiet.log
# racecheck, smallest case: B=64 (two CTAs), T=9, with z/c stores
PROBE_T=9 timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all $P $L fresh 64 0 1 0 > $O/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 15 $O/racecheck.log
