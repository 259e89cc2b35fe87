#!/bin/bash
# round-2 GPU session Q (one B200): suite with the empty-shard group test; sequential Float32 step kernel unrolled
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02q_pytest.log
tail -6 gpurun_out/r02q_pytest.log
OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C5s C5s C3 > gpurun_out/r02q_configs.jsonl 2> gpurun_out/r02q_err.log
python - <<'PY'
import json
for l in open('gpurun_out/r02q_configs.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:40], round(d['iterations_per_s'],1), d.get('kernels_ms_per_iteration'), d.get('kernel_launches_per_iteration'))
PY
