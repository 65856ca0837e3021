#!/bin/bash
# One GPU session: parity tests, headline bench (+ reference arm), merge bench, ncu launch list and one full capture.
# Usage (from the repo root on the GPU box): bash scripts/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 600 $OUT/${TAG}_bench.json
timeout 300 python bench.py --workload merge --steps 3 --warmup 3 > $OUT/${TAG}_bench_merge.json 2> $OUT/${TAG}_bench_merge.err; tail -c 900 $OUT/${TAG}_bench_merge.json
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
timeout 300 python scripts/gpu_phases.py > $OUT/${TAG}_phases.log 2>&1; tail -20 $OUT/${TAG}_phases.log
DG_WORKLOAD=merge timeout 300 python scripts/gpu_phases.py 4096 > $OUT/${TAG}_phases_merge.log 2>&1; tail -20 $OUT/${TAG}_phases_merge.log
for W in agents3 agents4 curve45 curve90; do DG_WORKLOAD=$W timeout 200 python scripts/gpu_phases.py 1480 > $OUT/${TAG}_phases_$W.log 2>&1; head -2 $OUT/${TAG}_phases_$W.log; done
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -3 $OUT/${TAG}_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --batch 1480 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
DG_WORKLOAD=merge timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgsqp_solve_kernel -c 1 -o $OUT/${TAG}_merge_full -f python scripts/profile_small.py 148 > $OUT/${TAG}_ncu_merge_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgsqp_solve_kernel -c 1 -o $OUT/${TAG}_solve_full -f python scripts/profile_small.py 148 > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -20
