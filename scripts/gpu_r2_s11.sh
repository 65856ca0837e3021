#!/bin/bash
# Round 2, GPU session 11: full GPU test suite (new round-2 parity tests) + multi-agent phases after the L2 tridiag change.
OUT=gpurun_out; TAG=${TAG:-r2_s11}
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -6 $OUT/${TAG}_pytest_gpu.log
export DGSQP_B200_LIB=$PWD/dgsqp_b200/libdgsqp_b200_prof.so
for W in agents3 agents4; do DG_WORKLOAD=$W timeout 300 python scripts/gpu_phases.py 592 > $OUT/${TAG}_phases_$W.log 2>&1; grep -v "^  " $OUT/${TAG}_phases_$W.log | head -1; grep "pd_tridiag\|cholesky\|tri_inverse\|total mean" $OUT/${TAG}_phases_$W.log; done
