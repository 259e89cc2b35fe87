#!/bin/bash
# round-2 GPU session N (one B200): validation of the final defaults: suite, smoke, bench, launch list, ncu capture
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02n_pytest.log
tail -6 gpurun_out/r02n_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02n_bench_n1.json 2> gpurun_out/r02n_bench_n1.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02n_bench_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-other-configs > /dev/null 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_direction2 -s 16 -c 1 -f -o gpurun_out/r02n_direction python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-other-configs --no-e2e > gpurun_out/r02n_ncu1.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02n_bench_n1.json') if l.startswith('{')][0])
print(round(d['value'],2), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), {k:(v['launches'],round(v['avg_ms'],3)) for k,v in d['roofline']['kernels'].items()}, d['clocks'], d['e2e']['value'], d['roofline']['traffic'])
print({k:round(v.get('iterations_per_s',0),1) for k,v in d['other_configs'].items()})
print(d['cpu_baseline']['value'], d['cpu_baseline']['single_thread']['value'])
PY
