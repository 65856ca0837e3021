#!/bin/bash
# Round 2, GPU session 3: blocked warm start (in-place D, panel QR + compact-WY apply): parity, phases, speed.
OUT=gpurun_out; TAG=${TAG:-r2_s3}
mkdir -p $OUT
set -x
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -25 $OUT/${TAG}_pytest_gpu.log
timeout 300 python scripts/gpu_parity1000.py > $OUT/${TAG}_parity1000.log 2>&1; cat $OUT/${TAG}_parity1000.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 1500 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
timeout 600 python bench.py --workload merge --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_merge.json 2> $OUT/${TAG}_bench_merge.err; tail -c 1200 $OUT/${TAG}_bench_merge.json; tail -3 $OUT/${TAG}_bench_merge.err
export DGSQP_B200_LIB=$PWD/dgsqp_b200/libdgsqp_b200_prof.so
timeout 300 python scripts/gpu_phases.py 2048 > $OUT/${TAG}_phases_chicane.log 2>&1; tail -32 $OUT/${TAG}_phases_chicane.log
DG_WORKLOAD=merge timeout 300 python scripts/gpu_phases.py 4096 > $OUT/${TAG}_phases_merge.log 2>&1; tail -32 $OUT/${TAG}_phases_merge.log
