#!/bin/bash
# gpurun driver: GPU suite, default bench line, ncu launch list of the same command, ncu --set full of the dominant kernels
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/plain_short.json 2> $O/plain_short.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 200 python tools/prof_wide.py 32 128 8192 > $O/plain_c2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_fwd_kernel -s 2 -c 1 -o $O/prof_c2 -f python tools/prof_wide.py 32 128 8192 > $O/ncu_c2.log 2>&1
timeout 200 python tools/prof_wide.py 64 256 8192 > $O/plain_m1.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_ -s 2 -c 2 -o $O/prof_m1 -f python tools/prof_wide.py 64 256 8192 > $O/ncu_m1.log 2>&1
timeout 200 python tools/prof_wide.py 256 128 8192 > $O/plain_m2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_ -s 2 -c 2 -o $O/prof_m2 -f python tools/prof_wide.py 256 128 8192 > $O/ncu_m2.log 2>&1
ls -la $O | head -30
