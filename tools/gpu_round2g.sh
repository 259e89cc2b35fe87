#!/bin/bash
# round-2 GPU session G (one B200): validation of the running-share give-up and the Float32 carry default
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02g_pytest.log
tail -8 gpurun_out/r02g_pytest.log
OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C5s-gram C2 C4 > gpurun_out/r02g_configs_native.jsonl 2> gpurun_out/r02g_err.log
cat gpurun_out/r02g_configs_native.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_direction2|k_gram2|k_trial" -c 70 --csv --log-file gpurun_out/r02g_c5s_launches.csv python tools/bench_configs.py C5s-gram > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_bench_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-other-configs > /dev/null 2>&1
