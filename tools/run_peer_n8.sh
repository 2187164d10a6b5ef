# eight GPUs: latency of the gradient exchange in isolation, the multi-GPU correctness check, the data-parallel training step
N=${1:-8}
O=gpurun_out/peer$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 100 $TR --master-port 29537 tools/time_collective.py > $O/time_collective.log 2>&1; echo "time_collective rc=$?"; grep -E "COLLECTIVE|Error|error" $O/time_collective.log | tail -3
timeout 150 $TR --master-port 29533 tests/dist_check.py > $O/dist_check.log 2>&1; echo "dist_check rc=$?"; grep -E "DIST_CHECK|Error|error|assert" $O/dist_check.log | tail -5
Q="--workload c3 --steps 50 --no-extra --no-cpu-baseline --no-e2e"
for c in peer nccl; do
  timeout 120 $TR --master-port 29541 bench.py --gpus $N $Q --collective $c > $O/bench_$c.json 2> $O/bench_$c.err; echo "bench $c rc=$?"
  python -c "import sys,json; d=json.loads(open('$O/bench_$c.json').read().strip().splitlines()[-1]); print('$c', d['impl_detail'].get('collective'), round(d['ms_per_step'],4), [round(v,4) for v in d.get('rank_ms_per_step')])" 2>&1 | tail -1
done
