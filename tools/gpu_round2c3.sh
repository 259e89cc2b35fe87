#!/bin/bash
# round-2 GPU session C3 (one B200): where the chained EXACT instance loses 8 us (ncu launch durations, debug switches)
mkdir -p gpurun_out
for V in "0 0" "1 0" "1 4" "1 6" "1 2"; do
set -- $V
OPKB_CHAIN=$1 OPKB_CHAIN_DBG=$2 OPKB_BENCH_NOPROF=1 OPKB_BENCH_NATIVE=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02c3_chain$1_dbg$2.csv python tools/bench_configs.py C1 > /dev/null 2>&1
done
python - <<'PY'
import csv,glob,statistics
for f in sorted(glob.glob('gpurun_out/r02c3_chain*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>5 and 'k_direction2' in r[4]]
    d=[float(r[-1].replace(',',''))/1e3 for r in rows]; big=[x for x in d if x>10]
    print(f[-18:], 'real', len(big), 'of', len(d), 'median', round(statistics.median(big),1), 'first', [round(x,1) for x in big[:8]], 'last', [round(x,1) for x in big[-4:]])
PY
