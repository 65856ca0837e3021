#!/bin/bash
# Round 2, GPU session 1: sanity (pytest -m gpu), the rewritten bench on every workload (short), fine-grained phase profiles.
OUT=gpurun_out; TAG=r2_s1
mkdir -p $OUT
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $OUT/${TAG}_gpu.txt
nproc >> $OUT/${TAG}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 1500 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; tail -c 600 $OUT/${TAG}_bench_ref.json
for W in merge curve agents3 agents4; do
  timeout 600 python bench.py --workload $W --steps 1 --warmup 3 > $OUT/${TAG}_bench_$W.json 2> $OUT/${TAG}_bench_$W.err; tail -c 700 $OUT/${TAG}_bench_$W.json; tail -3 $OUT/${TAG}_bench_$W.err
done
export DGSQP_B200_LIB=$PWD/dgsqp_b200/libdgsqp_b200_prof.so
timeout 300 python scripts/gpu_phases.py 2048 > $OUT/${TAG}_phases_chicane.log 2>&1; tail -32 $OUT/${TAG}_phases_chicane.log
DG_WORKLOAD=merge timeout 300 python scripts/gpu_phases.py 4096 > $OUT/${TAG}_phases_merge.log 2>&1; tail -32 $OUT/${TAG}_phases_merge.log
for W in agents3 agents4; do DG_WORKLOAD=$W timeout 400 python scripts/gpu_phases.py 592 > $OUT/${TAG}_phases_$W.log 2>&1; tail -30 $OUT/${TAG}_phases_$W.log; done
ls -la $OUT | tail -20
