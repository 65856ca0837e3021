#!/bin/bash
# Quick GPU check of an eigen-stage change: stage + golden + round-2 parity tests, nearestPD-heavy cases, 1000-instance parity,
# chicane + merge bench (short), chicane phases.
OUT=gpurun_out; TAG=${TAG:-quick3}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stages or golden or live_oracle or round2 or small_horizon or multi_agent" > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python scripts/gpu_parity1000.py 2>&1 | head -2 | cut -c1-200
for W in chicane merge; do timeout 900 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$W.json 2> $OUT/${TAG}_bench_$W.err; python -c "
import json
for l in open('$OUT/${TAG}_bench_$W.json'):
    if l.startswith('{'): d=json.loads(l); print('$W', d['value'], d['solves_per_sec_all'], d['e2e']['value'])
"; done
export DGSQP_B200_LIB=$PWD/dgsqp_b200/libdgsqp_b200_prof.so
timeout 300 python scripts/gpu_phases.py 2048 > $OUT/${TAG}_phases_chicane.log 2>&1; grep "pd_eigval\|pd_invit\|total mean" $OUT/${TAG}_phases_chicane.log
