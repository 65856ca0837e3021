#!/bin/bash
# Round 2, GPU session 4: tests + bench lines + ncu full capture (source-level) of the blocked warm-start build.
OUT=gpurun_out; TAG=${TAG:-r2_s4}
mkdir -p $OUT
set -x
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 700 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
timeout 600 python bench.py --workload merge --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_merge.json 2> $OUT/${TAG}_bench_merge.err; tail -c 600 $OUT/${TAG}_bench_merge.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgsqp_solve_kernel -c 1 -o $OUT/${TAG}_solve_full -f python scripts/profile_small.py 148 > $OUT/${TAG}_ncu_full.log 2>&1; tail -3 $OUT/${TAG}_ncu_full.log
ls -la $OUT | tail -6
