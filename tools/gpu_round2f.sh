#!/bin/bash
# round-2 GPU session F (one B200): carry give-up threshold; bench with the other configurations
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02f_pytest.log
tail -6 gpurun_out/r02f_pytest.log
OPKB_CARRY=2 OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C5s-gram > gpurun_out/r02f_c5s_gram_carry.jsonl 2> gpurun_out/r02f_err.log
cat gpurun_out/r02f_c5s_gram_carry.jsonl
OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C2 C4 > gpurun_out/r02f_configs_native.jsonl 2>> gpurun_out/r02f_err.log
cat gpurun_out/r02f_configs_native.jsonl
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02f_bench_n1.err
