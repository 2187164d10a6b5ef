#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/n2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 150 $TR bench.py --gpus 2 --steps 50 --warmup 3 --workload c3 --no-extra > gpurun_out/n2/c3_onegraph.json 2> gpurun_out/n2/c3_onegraph.err; echo "one-graph rc=$?"
FGRNN_BENCH_DP_GRAPH=two timeout 150 $TR bench.py --gpus 2 --steps 50 --warmup 3 --workload c3 --no-extra > gpurun_out/n2/c3_twograph.json 2> gpurun_out/n2/c3_twograph.err; echo "two-graph rc=$?"
timeout 240 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/n2/c2_full.json 2> gpurun_out/n2/c2_full.err; echo "c2 rc=$?"
tail -n 3 gpurun_out/n2/*.err
python - <<PY
import json
for f in ("c3_onegraph", "c3_twograph", "c2_full"):
    try:
        d = json.loads(open("gpurun_out/n2/%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in ("value", "ms_per_step", "rank_ms_per_step", "gpu_launches")}, d["impl_detail"], d["clocks"])
        if d.get("e2e"): print("  e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], (d["e2e"].get("last_state_only") or {}).get("value"))
        for k, v in (d.get("extra") or {}).items(): print("  ", k, v.get("value"), v.get("ms_per_step"), v.get("error"))
    except Exception as e:
        print(f, "no result:", e)
PY
