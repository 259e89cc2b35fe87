#!/bin/bash
# round-2 GPU session I (one B200): barrier ops with suspend hint; history A/B again
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02i_pytest.log
tail -6 gpurun_out/r02i_pytest.log
for H in 1 0; do
OPKB_HISTORY=$H python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-e2e > gpurun_out/r02i_bench_h$H.json 2> gpurun_out/r02i_bench_h$H.err; echo "bench h$H rc=$?"
OPKB_HISTORY=$H OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C2 C4 C5s-gram > gpurun_out/r02i_configs_h$H.jsonl 2>> gpurun_out/r02i_err.log
done
for H in 1 0; do
OPKB_HISTORY=$H python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-other-configs --no-e2e > gpurun_out/r02i_bench40_h$H.json 2> gpurun_out/r02i_bench40_h$H.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02i_bench*_h*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],2), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), {k:(v['launches'],round(v['avg_ms'],3)) for k,v in d['roofline']['kernels'].items()}, d['clocks']['sm_mhz'])
for f in sorted(glob.glob('gpurun_out/r02i_configs_h*.jsonl')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f[-20:], d['config'][:28], round(d['iterations_per_s'],1), d['kernels_ms_per_iteration'])
PY
