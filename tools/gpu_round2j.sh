#!/bin/bash
# round-2 GPU session J (one B200): SPEC instances of the fused kernel (compile-time flags, partial-free body): A/B, ncu
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02j_pytest.log
tail -6 gpurun_out/r02j_pytest.log
for S in 1 0 1 0; do
OPKB_DIR_SPEC=$S python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-other-configs --no-e2e >> gpurun_out/r02j_bench40_spec$S.json 2>> gpurun_out/r02j_err.log
done
for S in 1 0; do
OPKB_DIR_SPEC=$S OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C2 C4 > gpurun_out/r02j_configs_spec$S.jsonl 2>> gpurun_out/r02j_err.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02j_bench40_spec*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f[-28:], round(d['value'],2), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), {k:(v['launches'],round(v['avg_ms'],3)) for k,v in d['roofline']['kernels'].items()}, d['clocks']['sm_mhz'])
for f in sorted(glob.glob('gpurun_out/r02j_configs_spec*.jsonl')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f[-22:], d['config'][:28], round(d['iterations_per_s'],1), d['kernels_ms_per_iteration'])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_direction2 -s 16 -c 1 -f -o gpurun_out/r02j_direction_spec python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-other-configs --no-e2e > gpurun_out/r02j_ncu1.log 2>&1; echo "ncu rc=$?"
