#!/bin/bash
# Round 2, GPU session 2: validate warm start + polish + register Cholesky + Sturm rewrite (parity, race check, speed, phases).
OUT=gpurun_out; TAG=r2_s2
mkdir -p $OUT
set -x
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -25 $OUT/${TAG}_pytest_gpu.log
timeout 300 python scripts/gpu_parity1000.py > $OUT/${TAG}_parity1000.log 2>&1; cat $OUT/${TAG}_parity1000.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/gpu_race.py 0,1 3 > $OUT/${TAG}_racecheck.log 2>&1; tail -8 $OUT/${TAG}_racecheck.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 1800 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
DGSQP_QP_WARM=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_cold.json 2> $OUT/${TAG}_bench_cold.err; tail -c 900 $OUT/${TAG}_bench_cold.json
timeout 600 python bench.py --workload merge --steps 2 --warmup 3 > $OUT/${TAG}_bench_merge.json 2> $OUT/${TAG}_bench_merge.err; tail -c 1500 $OUT/${TAG}_bench_merge.json; tail -3 $OUT/${TAG}_bench_merge.err
export DGSQP_B200_LIB=$PWD/dgsqp_b200/libdgsqp_b200_prof.so
timeout 300 python scripts/gpu_phases.py 2048 > $OUT/${TAG}_phases_chicane.log 2>&1; tail -32 $OUT/${TAG}_phases_chicane.log
DG_WORKLOAD=merge timeout 300 python scripts/gpu_phases.py 4096 > $OUT/${TAG}_phases_merge.log 2>&1; tail -32 $OUT/${TAG}_phases_merge.log
DG_WORKLOAD=curve90 timeout 300 python scripts/gpu_phases.py 1480 > $OUT/${TAG}_phases_curve90.log 2>&1; tail -32 $OUT/${TAG}_phases_curve90.log
ls -la $OUT | tail -12
