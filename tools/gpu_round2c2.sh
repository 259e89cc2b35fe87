#!/bin/bash
# round-2 GPU session C2 (one B200): chained launches with the warp-cooperative decision and the L-BFGS instances -- tests, A/B
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -30 ) > gpurun_out/r02c2_chain_tests.log; tail -4 gpurun_out/r02c2_chain_tests.log
: > gpurun_out/r02c2_ab.txt
for CH in 0 1 0 1; do
OPKB_BENCH_NOPROF=1 OPKB_CHAIN=$CH OPKB_BENCH_NATIVE=1 timeout 300 python tools/bench_configs.py C1 C2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('chain=$CH', d['config'][:12], round(d['iterations_per_s'],1), d.get('chain_heads_adopted_cancelled'), d['f'])
" | tee -a gpurun_out/r02c2_ab.txt
done
OPKB_CHAIN=1 OPKB_BENCH_NOPROF=1 OPKB_BENCH_NATIVE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r02c2_c1_chain1_launches.csv python tools/bench_configs.py C1 > /dev/null 2>&1
python - <<'PY'
import csv,statistics
rows=[r for r in csv.reader(open('gpurun_out/r02c2_c1_chain1_launches.csv')) if len(r)>5 and 'k_direction2' in r[4]]
d=[float(r[-1].replace(',',''))/1e3 for r in rows]; big=[x for x in d if x>10]
print('chain kernel under ncu: n', len(big), 'median us', round(statistics.median(big),1), set(r[4][:62] for r in rows))
PY
