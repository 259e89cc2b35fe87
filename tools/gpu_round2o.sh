#!/bin/bash
# round-2 GPU session O (one B200): history without shared-memory stores: parity subset + A/B
mkdir -p gpurun_out
OPKB_HISTORY=2 python -m pytest tests/test_gpu_baseline_parity.py tests/test_gpu_carry.py tests/test_gpu_vmlmb.py -q > gpurun_out/r02o_pytest_hist2.log 2>&1; echo "pytest(hist=2) rc=$?" | tee -a gpurun_out/r02o_pytest_hist2.log
tail -3 gpurun_out/r02o_pytest_hist2.log
python -m pytest tests/test_gpu_baseline_parity.py tests/test_gpu_carry.py tests/test_gpu_vmlmb.py tests/test_gpu_group.py -q > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02o_pytest.log
tail -3 gpurun_out/r02o_pytest.log
for H in 2 1 2 1; do
OPKB_HISTORY=$H python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-other-configs --no-e2e >> gpurun_out/r02o_bench40_hist$H.json 2>> gpurun_out/r02o_err.log
done
for H in 1 0; do
OPKB_HISTORY=$H OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C2 C1 >> gpurun_out/r02o_configs_hist$H.jsonl 2>> gpurun_out/r02o_err.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02o_bench40_hist*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f[-28:], round(d['value'],2), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), {k:(v['launches'],round(v['avg_ms'],3)) for k,v in d['roofline']['kernels'].items()}, d['clocks']['sm_mhz'])
for f in sorted(glob.glob('gpurun_out/r02o_configs_hist*.jsonl')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f[-22:], d['config'][:28], round(d['iterations_per_s'],1), d['kernels_ms_per_iteration'])
PY
