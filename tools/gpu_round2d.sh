#!/bin/bash
# round-2 GPU session D (8 B200 of one box): hardware parity at 2 / 4 / 8 ranks, one process on 8 GPUs, bench lines
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | wc -l
( time python -m pytest tests/test_gpu_baseline_parity.py::test_two_ranks_on_hardware tests/test_gpu_group.py tests/test_gpu_api.py -q ) > gpurun_out/r02d_pytest_multi.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02d_pytest_multi.log
tail -4 gpurun_out/r02d_pytest_multi.log
for N in 4 8; do
  $TR --nproc-per-node $N --master-port 2951$N tools/multigpu_parity.py > gpurun_out/r02d_multigpu_parity_n$N.txt 2>&1; echo "parity N=$N rc=$?" | tee -a gpurun_out/r02d_multigpu_parity_n$N.txt
  grep -c " OK" gpurun_out/r02d_multigpu_parity_n$N.txt; grep "FAIL" gpurun_out/r02d_multigpu_parity_n$N.txt | head -3
done
$TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02d_bench_n8.json 2> gpurun_out/r02d_bench_n8.err; echo "bench8 rc=$?"
$TR --nproc-per-node 2 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err; echo "bench2 rc=$?"
python tools/bench_group.py 8 20 > gpurun_out/r02d_group_n8.jsonl 2> gpurun_out/r02d_group_n8.err; echo "group8 rc=$?"; cat gpurun_out/r02d_group_n8.jsonl
OPKB_BENCH_NATIVE=1 $TR --nproc-per-node 8 --master-port 29533 tools/bench_configs.py C5 C5-gram > gpurun_out/r02d_c5_n8.jsonl 2> gpurun_out/r02d_c5_n8.err; echo "c5 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02d_*.json*')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f, d.get('value') or d.get('iterations_per_s'), d.get('ms_per_step') or d.get('ms_per_iteration'), (d.get('e2e') or {}).get('value'), d.get('frac_of_measured_peak'))
PY
