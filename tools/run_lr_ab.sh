# A/B of the low-rank tensor-core kernel builds (kws_b200/lib/libfastgrnn_b200_v?.so), one box
# variants: make -C kws_b200/csrc variant TAG=vA DEFS="-DTL_HOP_HALVES=1 -DTL_RCP_NEWTON=0 -DTL_TMA_STORE=0"; vB: ...TMA_STORE=1; vC: + -DTL_RCP_NEWTON=1; vD: -DTL_HOP_HALVES=2 -DTL_TMA_STORE=1
L=/root/repo/kws_b200/lib
for v in B D; do KWS_B200_LIB=$L/libfastgrnn_b200_v$v.so timeout 200 python -m pytest tests/test_gpu_tc_lowrank.py -m gpu -x -q 2>&1 | tail -1; done
for rep in 1 2; do for v in A B C D; do echo "-- $v"; KWS_B200_LIB=$L/libfastgrnn_b200_v$v.so timeout 100 python tools/prof_lowrank.py 32768 99 30; done; done
