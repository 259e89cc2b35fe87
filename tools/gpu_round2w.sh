#!/bin/bash
# round-2 GPU session W (one B200): SPEC = 3 instances (L-BFGS without bounds) -- parity tests that run them, C1 A/B
mkdir -p gpurun_out
( timeout 800 python -m pytest tests/test_gpu_vmlmb.py tests/test_gpu_carry.py tests/test_gpu_chain.py tests/test_gpu_api.py "tests/test_gpu_baseline_parity.py::test_c1_unconstrained_rosenbrock_1e6" -x -q 2>&1 | tail -15 ) > gpurun_out/r02w_tests.log; tail -5 gpurun_out/r02w_tests.log
for SP in 0 1; do
OPKB_DIR_SPEC=$SP OPKB_BENCH_NATIVE=1 timeout 300 python tools/bench_configs.py C1 > gpurun_out/r02w_c1_spec$SP.jsonl 2> gpurun_out/r02w_err$SP.log
OPKB_DIR_SPEC=$SP OPKB_BENCH_NOPROF=1 OPKB_BENCH_NATIVE=1 timeout 300 python tools/bench_configs.py C1 >> gpurun_out/r02w_c1_spec$SP.jsonl 2>> gpurun_out/r02w_err$SP.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02w_c1_spec*.jsonl')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f[-12:], d['config'][:30], round(d['iterations_per_s'],1), d.get('kernels_ms_per_iteration'), d['f'])
PY
