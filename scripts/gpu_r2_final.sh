#!/bin/bash
# Round 2, final GPU session: full GPU test suite, smoke, bench lines of every BASELINE workload, the reference arm, the ncu
# launch list of the bench command, ncu --set full captures (chicane, merge) for the DRAM traffic and the source hotspots,
# phase profiles of every workload (profiling build).
OUT=gpurun_out; TAG=${TAG:-r2_final}
mkdir -p $OUT
set -x
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 400 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; tail -c 600 $OUT/${TAG}_bench_ref.json
for W in merge curve agents3 agents4; do timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > $OUT/${TAG}_bench_$W.json 2> $OUT/${TAG}_bench_$W.err; tail -c 300 $OUT/${TAG}_bench_$W.json; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --batch 1480 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgsqp_solve_kernel -c 1 -o $OUT/${TAG}_solve_full -f python scripts/profile_small.py 148 > $OUT/${TAG}_ncu_full.log 2>&1; tail -2 $OUT/${TAG}_ncu_full.log
DG_WORKLOAD=merge timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgsqp_solve_kernel -c 1 -o $OUT/${TAG}_merge_full -f python scripts/profile_small.py 148 > $OUT/${TAG}_ncu_merge_full.log 2>&1; tail -2 $OUT/${TAG}_ncu_merge_full.log
export DGSQP_B200_LIB=$PWD/dgsqp_b200/libdgsqp_b200_prof.so
timeout 300 python scripts/gpu_phases.py 2048 > $OUT/${TAG}_phases_chicane.log 2>&1
for W in merge curve45 curve90 agents3 agents4; do DG_WORKLOAD=$W timeout 300 python scripts/gpu_phases.py 1480 > $OUT/${TAG}_phases_$W.log 2>&1; head -1 $OUT/${TAG}_phases_$W.log; done
ls -la $OUT | tail -30
