#!/bin/bash
# Quick GPU check of a kernel change: stage-level + golden parity tests, 1000-instance parity, phase profile (chicane, merge).
OUT=gpurun_out; TAG=${TAG:-quick}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stages or golden or live_oracle or merge_vs" > $OUT/${TAG}_pytest.log 2>&1; tail -4 $OUT/${TAG}_pytest.log
timeout 300 python scripts/gpu_parity1000.py 2>&1 | tail -6 | cut -c1-260
export DGSQP_B200_LIB=$PWD/dgsqp_b200/libdgsqp_b200_prof.so
timeout 300 python scripts/gpu_phases.py 2048 > $OUT/${TAG}_phases_chicane.log 2>&1; tail -36 $OUT/${TAG}_phases_chicane.log
DG_WORKLOAD=merge timeout 300 python scripts/gpu_phases.py 2048 > $OUT/${TAG}_phases_merge.log 2>&1; tail -36 $OUT/${TAG}_phases_merge.log
