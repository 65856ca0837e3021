#!/bin/bash
# Round 2, GPU session 12: full GPU test suite, 1000-instance parity, phases (chicane, merge) after the direct sym load,
# H-build prefetch and two-sided Sturm count.
OUT=gpurun_out; TAG=${TAG:-r2_s12}
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -6 $OUT/${TAG}_pytest_gpu.log
timeout 300 python scripts/gpu_parity1000.py > $OUT/${TAG}_parity1000.log 2>&1; cut -c1-250 $OUT/${TAG}_parity1000.log
export DGSQP_B200_LIB=$PWD/dgsqp_b200/libdgsqp_b200_prof.so
timeout 300 python scripts/gpu_phases.py 2048 > $OUT/${TAG}_phases_chicane.log 2>&1; tail -30 $OUT/${TAG}_phases_chicane.log
DG_WORKLOAD=merge timeout 300 python scripts/gpu_phases.py 2048 > $OUT/${TAG}_phases_merge.log 2>&1; tail -3 $OUT/${TAG}_phases_merge.log | cut -c1-250
