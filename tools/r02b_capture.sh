mkdir -p gpurun_out/r02b; O=gpurun_out/r02b
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_lr_fwd -c 1 -s 3 -o $O/prof_c4_tc_lr python tools/prof_lowrank.py 32768 99 1 > $O/ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_bwd_fused -c 1 -s 5 -o $O/prof_c3_bwd_fused python bench.py --workload c3 --steps 4 --no-extra --no-cpu-baseline --no-e2e --no-graph > $O/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
ls -la $O
