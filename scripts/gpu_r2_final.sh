#!/bin/bash
# Round 2, final GPU sessions (split so that gpurun_out stays small; ncu reports are summarised on the box and deleted).
#   bash scripts/gpu_r2_final.sh bench | ncu | phases
OUT=gpurun_out; TAG=${TAG:-r2_final}; WHAT=${1:-bench}
mkdir -p $OUT
set -x
if [ "$WHAT" = bench ]; then
  timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -5 $OUT/${TAG}_pytest_gpu.log
  timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
  timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 300 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; tail -c 400 $OUT/${TAG}_bench_ref.json
  for W in merge curve agents3 agents4; do timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > $OUT/${TAG}_bench_$W.json 2> $OUT/${TAG}_bench_$W.err; tail -c 200 $OUT/${TAG}_bench_$W.json; done
fi
if [ "$WHAT" = ncu ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --batch 1480 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
  for W in chicane merge; do
    DG_WORKLOAD=$W timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgsqp_solve_kernel -c 1 -o /tmp/${TAG}_$W -f python scripts/profile_small.py 148 > $OUT/${TAG}_ncu_${W}.log 2>&1; tail -2 $OUT/${TAG}_ncu_${W}.log
    ncu -i /tmp/${TAG}_$W.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_${W}_raw.csv 2>/dev/null
    ncu -i /tmp/${TAG}_$W.ncu-rep --page source --print-source cuda,sass --csv > /tmp/${TAG}_$W_src.csv 2>/dev/null
    python scripts/ncu_hotspots.py /tmp/${TAG}_$W_src.csv 60 > $OUT/${TAG}_${W}_source_hotspots.md; head -30 $OUT/${TAG}_${W}_source_hotspots.md
  done
fi
if [ "$WHAT" = phases ]; then
  export DGSQP_B200_LIB=$PWD/dgsqp_b200/libdgsqp_b200_prof.so
  timeout 300 python scripts/gpu_phases.py 2048 > $OUT/${TAG}_phases_chicane.log 2>&1
  for W in merge curve45 curve90 agents3 agents4; do DG_WORKLOAD=$W timeout 400 python scripts/gpu_phases.py 1480 > $OUT/${TAG}_phases_$W.log 2>&1; head -1 $OUT/${TAG}_phases_$W.log; done
fi
ls -la $OUT | tail -24
