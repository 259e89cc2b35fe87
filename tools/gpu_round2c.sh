#!/bin/bash
# round-2 GPU session C (one B200): full GPU suite; the Float32 carry kernel with / without SPLIT: per-launch times, ncu
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02c_pytest.log
tail -15 gpurun_out/r02c_pytest.log
OPKB_CARRY=2 OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C5s-gram > gpurun_out/r02c_c5s_gram_carry_split.jsonl 2> gpurun_out/r02c_err.log
OPKB_CARRY=2 OPKB_DIR_SPLIT=0 OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C5s-gram > gpurun_out/r02c_c5s_gram_carry_nosplit.jsonl 2>> gpurun_out/r02c_err.log
cat gpurun_out/r02c_c5s_gram_carry_split.jsonl gpurun_out/r02c_c5s_gram_carry_nosplit.jsonl
OPKB_CARRY=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_direction2 -c 40 --csv --log-file gpurun_out/r02c_dir2_split_launches.csv python tools/bench_configs.py C5s-gram > /dev/null 2>&1
OPKB_CARRY=2 OPKB_DIR_SPLIT=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_direction2 -c 40 --csv --log-file gpurun_out/r02c_dir2_nosplit_launches.csv python tools/bench_configs.py C5s-gram > /dev/null 2>&1
OPKB_CARRY=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_direction2 -s 14 -c 1 -f -o gpurun_out/r02c_dir2_f32_split python tools/bench_configs.py C5s-gram > gpurun_out/r02c_ncu1.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | grep r02c
