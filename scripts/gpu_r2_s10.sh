#!/bin/bash
# Round 2, GPU session 10: hybrid tile tridiagonalisation for n > 128: multi-agent tests + phases of all workloads.
OUT=gpurun_out; TAG=${TAG:-r2_s10}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stages or golden or live_oracle or merge_vs or multi_agent" > $OUT/${TAG}_pytest.log 2>&1; tail -4 $OUT/${TAG}_pytest.log
export DGSQP_B200_LIB=$PWD/dgsqp_b200/libdgsqp_b200_prof.so
for W in agents3 agents4 curve45 curve90; do DG_WORKLOAD=$W timeout 300 python scripts/gpu_phases.py 592 > $OUT/${TAG}_phases_$W.log 2>&1; grep -v "^  " $OUT/${TAG}_phases_$W.log | head -3; grep "pd_tridiag\|cholesky\|tri_inverse\|gi_dz\|gi_add\|ws_apply\|total mean" $OUT/${TAG}_phases_$W.log; done
