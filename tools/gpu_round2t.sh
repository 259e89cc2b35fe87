#!/bin/bash
# round-2 GPU session T (one B200): device-resident decisions (chained launches) -- tests, C1/C2 A/B
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -40 ) > gpurun_out/r02t_chain_tests.log; tail -25 gpurun_out/r02t_chain_tests.log
for CH in 0 1; do
OPKB_CHAIN=$CH OPKB_BENCH_NATIVE=1 timeout 300 python tools/bench_configs.py C1 C2 > gpurun_out/r02t_configs_chain$CH.jsonl 2> gpurun_out/r02t_err$CH.log
OPKB_BENCH_NOPROF=1 OPKB_CHAIN=$CH OPKB_BENCH_NATIVE=1 timeout 300 python tools/bench_configs.py C1 C2 >> gpurun_out/r02t_configs_chain$CH.jsonl 2>> gpurun_out/r02t_err$CH.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02t_configs_chain*.jsonl')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f[-12:], d['config'][:40], round(d['iterations_per_s'],1), d.get('chain_heads_adopted_cancelled'), d.get('incremental_gram_hits_misses'), d.get('kernels_ms_per_iteration'))
PY
tail -3 gpurun_out/r02t_err1.log
