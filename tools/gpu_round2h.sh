#!/bin/bash
# round-2 GPU session H (one B200): iterate history (29-pass fused kernel): full suite, A/B bench, ncu
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02h_pytest.log
tail -12 gpurun_out/r02h_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > gpurun_out/r02h_bench_hist.json 2> gpurun_out/r02h_bench_hist.err; echo "bench hist rc=$?"
OPKB_HISTORY=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-e2e > gpurun_out/r02h_bench_nohist.json 2> gpurun_out/r02h_bench_nohist.err; echo "bench nohist rc=$?"
OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C1 C2 C4 C5s-gram > gpurun_out/r02h_configs_hist.jsonl 2> gpurun_out/r02h_err.log
OPKB_HISTORY=0 OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C1 C2 C4 C5s-gram > gpurun_out/r02h_configs_nohist.jsonl 2>> gpurun_out/r02h_err.log
python - <<'PY'
import json
for f in ('gpurun_out/r02h_bench_hist.json','gpurun_out/r02h_bench_nohist.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],2), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), {k:(v['launches'],round(v['avg_ms'],3)) for k,v in d['roofline']['kernels'].items()}, d['roofline'].get('iterate_history'))
for f in ('gpurun_out/r02h_configs_hist.jsonl','gpurun_out/r02h_configs_nohist.jsonl'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f[-20:], d['config'][:28], round(d['iterations_per_s'],1), d['kernels_ms_per_iteration'])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_direction2 -s 16 -c 1 -f -o gpurun_out/r02h_direction_hist python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-other-configs --no-e2e > gpurun_out/r02h_ncu1.log 2>&1; echo "ncu rc=$?"
