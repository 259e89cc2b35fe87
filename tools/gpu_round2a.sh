#!/bin/bash
# round-2 GPU session A (one B200): full GPU suite, both bench arms, config lines, ncu of the Float32 kernels
mkdir -p gpurun_out
nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; echo "bench rc=$?"
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r02a_bench_reference_n1.json 2> gpurun_out/r02a_bench_reference_n1.err; echo "ref rc=$?"
OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C1 C2 C3 C5s C5s-gram > gpurun_out/r02a_configs_native.jsonl 2> gpurun_out/r02a_configs_native.err
OPKB_CARRY=2 OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C5s-gram > gpurun_out/r02a_c5s_gram_carry.jsonl 2>> gpurun_out/r02a_configs_native.err
python tools/bench_configs.py C1 C2 > gpurun_out/r02a_configs_python.jsonl 2>> gpurun_out/r02a_configs_native.err
OPKB_CARRY=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_direction2 -s 14 -c 1 -f -o gpurun_out/r02a_dir2_f32 python tools/bench_configs.py C5s-gram > gpurun_out/r02a_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_twoloop_step -s 40 -c 2 -f -o gpurun_out/r02a_twoloop_f32 python tools/bench_configs.py C5s > gpurun_out/r02a_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spg -s 6 -c 1 -f -o gpurun_out/r02a_spg_f32 python tools/bench_configs.py C3 > gpurun_out/r02a_ncu3.log 2>&1; echo "ncu3 rc=$?"
ls -la gpurun_out | tail -20
