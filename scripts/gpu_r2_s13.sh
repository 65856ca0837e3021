#!/bin/bash
# Round 2, GPU session 13: full GPU suite + parity + phases + bench after: 2 inverse-iteration sweeps, shared-memory
# inverse-iteration scratch for the split placement, fused statistics epilogue, cross-chunk cluster fix, step_batch.
OUT=gpurun_out; TAG=${TAG:-r2_s13}
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 300 python scripts/gpu_parity1000.py > $OUT/${TAG}_parity1000.log 2>&1; cut -c1-250 $OUT/${TAG}_parity1000.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; python -c "
import json
for l in open('$OUT/${TAG}_bench.json'):
    if l.startswith('{'): d=json.loads(l); print('chicane', d['value'], d['solves_per_sec_all'], d['e2e']['value'])
"
timeout 900 python bench.py --workload merge --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_merge.json 2> $OUT/${TAG}_bench_merge.err; python -c "
import json
for l in open('$OUT/${TAG}_bench_merge.json'):
    if l.startswith('{'): d=json.loads(l); print('merge', d['value'], d['e2e']['value'])
"
export DGSQP_B200_LIB=$PWD/dgsqp_b200/libdgsqp_b200_prof.so
timeout 300 python scripts/gpu_phases.py 2048 > $OUT/${TAG}_phases_chicane.log 2>&1; grep "pd_invit\|pd_eigval\|total mean" $OUT/${TAG}_phases_chicane.log
DG_WORKLOAD=merge timeout 300 python scripts/gpu_phases.py 2048 > $OUT/${TAG}_phases_merge.log 2>&1; grep "pd_invit\|pd_eigval\|total mean" $OUT/${TAG}_phases_merge.log
