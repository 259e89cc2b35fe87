#!/bin/bash
# round-2 GPU session S (one B200): full suite on HEAD after the group fixes; smoke; the driver's bench command
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02s_pytest.log
tail -6 gpurun_out/r02s_pytest.log
python __graft_entry__.py smoke > gpurun_out/r02s_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02s_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02s_bench_n1.json 2> gpurun_out/r02s_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02s_bench_n1.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline'], d.get('clocks'))
PY
