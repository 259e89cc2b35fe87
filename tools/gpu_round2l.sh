#!/bin/bash
# round-2 GPU session L (one B200): Float32 iterate history with the SPEC instance: A/B + parity
mkdir -p gpurun_out
OPKB_HISTORY=2 python -m pytest tests/test_gpu_baseline_parity.py::test_c5_workload_float32 tests/test_gpu_carry.py tests/test_gpu_vmlmb.py -q > gpurun_out/r02l_pytest_hist2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02l_pytest_hist2.log
tail -4 gpurun_out/r02l_pytest_hist2.log
for H in 2 1 2 1; do
OPKB_HISTORY=$H OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C5s-gram >> gpurun_out/r02l_c5s_hist$H.jsonl 2>> gpurun_out/r02l_err.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02l_c5s_hist*.jsonl')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f[-22:], d['config'][:28], round(d['iterations_per_s'],1), d['kernels_ms_per_iteration'], d['incremental_gram_hits_misses'])
PY
