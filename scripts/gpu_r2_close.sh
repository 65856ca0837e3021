#!/bin/bash
# Round 2, closing GPU session on the final build: full GPU suite, smoke, bench lines (chicane default, reference arm, merge,
# curve, agents3), phases (chicane, merge), parity log.
OUT=gpurun_out; TAG=${TAG:-r2_close}
mkdir -p $OUT
set -x
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 300 python scripts/gpu_parity1000.py > $OUT/${TAG}_parity1000.log 2>&1; head -2 $OUT/${TAG}_parity1000.log | cut -c1-200
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -3 $OUT/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
for W in merge curve agents3; do timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > $OUT/${TAG}_bench_$W.json 2> $OUT/${TAG}_bench_$W.err; done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_close_bench*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], 'value %.1f'%d['value'], 'all %.1f'%d.get('solves_per_sec_all',0), 'e2e %.1f'%d['e2e']['value'], 'frac', d.get('roofline',{}).get('frac'))
P
export DGSQP_B200_LIB=$PWD/dgsqp_b200/libdgsqp_b200_prof.so
timeout 300 python scripts/gpu_phases.py 2048 > $OUT/${TAG}_phases_chicane.log 2>&1; tail -1 $OUT/${TAG}_phases_chicane.log | cut -c1-200
DG_WORKLOAD=merge timeout 300 python scripts/gpu_phases.py 2048 > $OUT/${TAG}_phases_merge.log 2>&1; tail -1 $OUT/${TAG}_phases_merge.log | cut -c1-200
