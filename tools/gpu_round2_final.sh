#!/bin/bash
# round-2 final GPU session (one B200): suite, smoke, the driver's bench command, reference arm (short), launch list,
# ncu --set full capture of the headline kernel and of C1's L-BFGS instance
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02f_pytest.log
tail -4 gpurun_out/r02f_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_bench_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-other-configs > /dev/null 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_direction2 -s 16 -c 1 -f -o gpurun_out/r02f_direction python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-other-configs --no-e2e > gpurun_out/r02f_ncu1.log 2>&1; echo "ncu rc=$?"
OPKB_BENCH_NOPROF=1 OPKB_BENCH_NATIVE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_direction2 -s 40 -c 1 -f -o gpurun_out/r02f_c1_direction python tools/bench_configs.py C1 > gpurun_out/r02f_ncu2.log 2>&1; echo "ncu C1 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02f_bench_n1.json') if l.startswith('{')][0])
print(round(d['value'],2), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), {k:(v['launches'],round(v['avg_ms'],3)) for k,v in d['roofline']['kernels'].items()}, d['clocks'], d['e2e']['value'], d['roofline']['traffic'])
print({k:round(v.get('iterations_per_s',0),1) for k,v in d['other_configs'].items()})
print(d['cpu_baseline']['value'], d['cpu_baseline']['single_thread']['value'])
PY
