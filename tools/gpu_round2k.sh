#!/bin/bash
# round-2 GPU session K (one B200): Float32 SPEC instance, integer counter: tests + config lines + bench
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02k_pytest.log
tail -6 gpurun_out/r02k_pytest.log
for S in 1 0; do
OPKB_DIR_SPEC=$S OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C5s-gram C2 C4 > gpurun_out/r02k_configs_spec$S.jsonl 2>> gpurun_out/r02k_err.log
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02k_configs_spec*.jsonl')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f[-22:], d['config'][:28], round(d['iterations_per_s'],1), d['kernels_ms_per_iteration'])
d=json.loads([l for l in open('gpurun_out/r02k_bench_n1.json') if l.startswith('{')][0])
print(round(d['value'],2), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), {k:(v['launches'],round(v['avg_ms'],3)) for k,v in d['roofline']['kernels'].items()}, d['clocks'], d['e2e']['value'])
print({k:round(v.get('iterations_per_s',0),1) for k,v in d['other_configs'].items()})
PY
