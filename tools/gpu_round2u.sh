#!/bin/bash
# round-2 GPU session U (one B200): where the chained kernel spends its extra time (ncu launch durations)
mkdir -p gpurun_out
for V in "0 0" "1 0" "1 1" "1 2"; do
set -- $V
OPKB_CHAIN=$1 OPKB_CHAIN_DBG=$2 OPKB_BENCH_NOPROF=1 OPKB_BENCH_NATIVE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r02u_c1_chain$1_dbg$2.csv python tools/bench_configs.py C1 > gpurun_out/r02u_c1_chain$1_dbg$2.log 2>&1
OPKB_CHAIN=$1 OPKB_CHAIN_DBG=$2 OPKB_BENCH_NOPROF=1 OPKB_BENCH_NATIVE=1 timeout 300 python tools/bench_configs.py C1 > gpurun_out/r02u_c1_chain$1_dbg$2.jsonl 2>&1
done
python - <<'PY'
import csv,glob,statistics,json
for f in sorted(glob.glob('gpurun_out/r02u_c1_chain*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>5 and 'k_direction2' in r[4]]
    d=[float(r[-1].replace(',','')) for r in rows[20:]]
    names=set(r[4][:60] for r in rows)
    print(f[-18:], len(d), 'median us', statistics.median(d)/1e3 if d else None, 'min', min(d)/1e3 if d else None, names)
for f in sorted(glob.glob('gpurun_out/r02u_c1_chain*.jsonl')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f[-20:], round(d['iterations_per_s'],1), d.get('chain_heads_adopted_cancelled'))
PY
