mkdir -p gpurun_out/r02c; O=gpurun_out/r02c
timeout 90 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
timeout 45 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 100 python bench.py --cpu-seconds 4 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; head -c 600 $O/bench_default.json
