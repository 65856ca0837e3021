#!/bin/bash
# Quick GPU check: stage + golden tests, 1000-instance parity, chicane bench (short), chicane phases.
OUT=gpurun_out; TAG=${TAG:-quick2}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stages or golden or live_oracle or round2" > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python scripts/gpu_parity1000.py 2>&1 | head -1
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; python -c "
import json
for l in open('$OUT/${TAG}_bench.json'):
    if l.startswith('{'): d=json.loads(l); print('chicane', d['value'], d['solves_per_sec_all'], d['e2e']['value'])
"
export DGSQP_B200_LIB=$PWD/dgsqp_b200/libdgsqp_b200_prof.so
timeout 300 python scripts/gpu_phases.py 2048 > $OUT/${TAG}_phases_chicane.log 2>&1; grep "hessian\|lin_\|adj_\|total mean" $OUT/${TAG}_phases_chicane.log
