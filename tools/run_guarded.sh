#!/bin/bash
# usage: tools/run_guarded.sh <seconds> <logfile> <command...>   -- runs the command in its own process group and
# kills the whole group (exact pgid, no pattern matching) if it is still alive after <seconds>
secs=$1; log=$2; shift 2
setsid bash -c "$*" > "$log" 2>&1 &
pid=$!
for i in $(seq 1 "$secs"); do sleep 1; kill -0 $pid 2>/dev/null || break; done
if kill -0 $pid 2>/dev/null; then echo "[run_guarded] still running after ${secs}s: killing process group $pid" >> "$log"; kill -KILL -- -$pid 2>/dev/null; sleep 2; fi
wait $pid 2>/dev/null
echo "[run_guarded] exit $?" >> "$log"
