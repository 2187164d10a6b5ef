# flagship forward: TMA tile stores (shipped lib) against the scalar-store build (libfastgrnn_b200_vS.so), and 64- / 56- / 48-row CTAs
# variant: make -C kws_b200/csrc variant TAG=vS DEFS="-DTC_TMA_STORE=0"   (TC_TMA_STORE=1: staging before the hand-off)
L=/root/repo/kws_b200/lib
O=gpurun_out/tcab; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
Q="--no-cpu-baseline --no-e2e --no-extra"
for rep in 1 2; do for v in "" _vS; do for vr in 16 14; do
  for w in "c2 100" "c5 8" "c4 10"; do set -- $w
    FGRNN_TC_VR=$vr KWS_B200_LIB=$L/libfastgrnn_b200$v.so timeout 200 python bench.py --workload $1 --steps $2 $Q 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib$v vr$vr', '$1', round(d['ms_per_step'],4), (d.get('roofline') or {}).get('frac'))"
  done; done; done; done
for v in "" _vS; do KWS_B200_LIB=$L/libfastgrnn_b200$v.so timeout 200 python bench.py --workload c3 --steps 40 $Q 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib$v', 'c3', round(d['ms_per_step'],4))"; done
