#!/bin/bash
# round-2 GPU session V (one B200): full suite with the chained-launch code in (opt-in), bench with the NUMA-near e2e buffers
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02v_pytest.log
tail -6 gpurun_out/r02v_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02v_bench_n1.json 2> gpurun_out/r02v_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02v_bench_n1.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['breakdown_s'], d['roofline']['frac'], d.get('clocks'))
        print({k:(v.get('iterations_per_s'),v.get('frac_of_measured_peak')) for k,v in d.get('other_configs',{}).items()})
PY
tail -3 gpurun_out/r02v_bench_n1.err
