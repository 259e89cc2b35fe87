#!/bin/bash
# round-2 GPU session Z (one B200): last-CTA combination 4 slots per warp and pass -- tests that depend on it, C1 / C2
mkdir -p gpurun_out
( timeout 800 python -m pytest tests/test_gpu_vmlmb.py tests/test_gpu_carry.py tests/test_gpu_chain.py tests/test_gpu_primitives.py tests/test_gpu_spg.py -x -q 2>&1 | tail -5 ) > gpurun_out/r02z_tests.log; tail -3 gpurun_out/r02z_tests.log
OPKB_BENCH_NATIVE=1 timeout 300 python tools/bench_configs.py C1 C2 > gpurun_out/r02z_configs.jsonl 2> gpurun_out/r02z_err.log
OPKB_BENCH_NOPROF=1 OPKB_BENCH_NATIVE=1 timeout 300 python tools/bench_configs.py C1 C2 >> gpurun_out/r02z_configs.jsonl 2>> gpurun_out/r02z_err.log
python - <<'PY'
import json
for l in open('gpurun_out/r02z_configs.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:30], round(d['iterations_per_s'],1), d.get('kernels_ms_per_iteration'), d['f'])
PY
