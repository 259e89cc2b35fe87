#!/bin/bash
# round-2 GPU session X (8 B200 of one box): final build -- multi-rank tests, parity at 8 ranks, bench at 8 and 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( timeout 300 python -m pytest tests/test_gpu_group.py "tests/test_gpu_baseline_parity.py::test_two_ranks_on_hardware" -q -x 2>&1 | tail -5 ) > gpurun_out/r02x_pytest_multi.log; tail -3 gpurun_out/r02x_pytest_multi.log
timeout 300 $TR --nproc-per-node 8 --master-port 29544 tools/multigpu_parity.py > gpurun_out/r02x_multigpu_parity_n8.txt 2>&1; echo "parity rc=$?"; grep -c " OK" gpurun_out/r02x_multigpu_parity_n8.txt
for N in 8 2; do
timeout 300 $TR --nproc-per-node $N --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02x_bench_n$N.json 2> gpurun_out/r02x_bench_n$N.err; echo "bench$N rc=$?"
done
OPKB_BENCH_NATIVE=1 timeout 300 $TR --nproc-per-node 8 --master-port 29543 tools/bench_configs.py C5 C5-gram > gpurun_out/r02x_c5_n8.jsonl 2> gpurun_out/r02x_c5_n8.err; echo "c5 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02x_*.json*')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f, round(d.get('value') or d.get('iterations_per_s'),2), round(d.get('ms_per_step') or d.get('ms_per_iteration'),4), (d.get('e2e') or {}).get('value'), d.get('frac_of_measured_peak'), (d.get('e2e') or {}).get('breakdown_s'), (d.get('clocks') or {}).get('sm_mhz'))
PY
