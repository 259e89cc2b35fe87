#!/bin/bash
# round-2 GPU session P (8 B200 of one box): final build -- bench at 8 / 4 / 2, C5 at 8, one process on 8 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4 2; do
$TR --nproc-per-node $N --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02p_bench_n$N.json 2> gpurun_out/r02p_bench_n$N.err; echo "bench$N rc=$?"
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > gpurun_out/r02p_bench_n1.json 2> gpurun_out/r02p_bench_n1.err; echo "bench1 rc=$?"
python tools/bench_group.py 8 20 > gpurun_out/r02p_group_n8.jsonl 2> gpurun_out/r02p_group_n8.err; echo "group8 rc=$?"
OPKB_BENCH_NATIVE=1 $TR --nproc-per-node 8 --master-port 29543 tools/bench_configs.py C5 C5-gram > gpurun_out/r02p_c5_n8.jsonl 2> gpurun_out/r02p_c5_n8.err; echo "c5 rc=$?"
$TR --nproc-per-node 8 --master-port 29544 tools/multigpu_parity.py > gpurun_out/r02p_multigpu_parity_n8.txt 2>&1; echo "parity rc=$?"; grep -c " OK" gpurun_out/r02p_multigpu_parity_n8.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02p_*.json*')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f, round(d.get('value') or d.get('iterations_per_s'),2), round(d.get('ms_per_step') or d.get('ms_per_iteration'),4), (d.get('e2e') or {}).get('value'), d.get('frac_of_measured_peak'), (d.get('e2e') or {}).get('breakdown_s'), (d.get('clocks') or {}).get('sm_mhz'))
PY
