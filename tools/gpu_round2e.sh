#!/bin/bash
# round-2 GPU session E (one B200): full GPU suite after the chain / Float32 Gram sweep changes; C5 shard lines; ncu
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02e_pytest.log
tail -12 gpurun_out/r02e_pytest.log
OPKB_CARRY=2 OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C5s-gram > gpurun_out/r02e_c5s_gram_carry.jsonl 2> gpurun_out/r02e_err.log
OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C5s-gram C5s C2 C1 C4 > gpurun_out/r02e_configs_native.jsonl 2>> gpurun_out/r02e_err.log
cat gpurun_out/r02e_c5s_gram_carry.jsonl gpurun_out/r02e_configs_native.jsonl
OPKB_CARRY=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_direction2|k_gram2|k_trial" -c 80 --csv --log-file gpurun_out/r02e_c5s_launches.csv python tools/bench_configs.py C5s-gram > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gram2 -s 12 -c 1 -f -o gpurun_out/r02e_gram2_f32 python tools/bench_configs.py C5s-gram > gpurun_out/r02e_ncu1.log 2>&1; echo "ncu rc=$?"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; echo "bench rc=$?"
ls -la gpurun_out | grep r02e
