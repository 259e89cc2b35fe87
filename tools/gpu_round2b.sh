#!/bin/bash
# round-2 GPU session B (one B200): the full GPU suite (no -x), clocks during the Float32 carry run, gram2 f32 ncu
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02b_pytest.log
tail -30 gpurun_out/r02b_pytest.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown --format=csv,noheader,nounits -lms 100 > gpurun_out/r02b_clocks_c5s.csv &
SMI=$!
sleep 1; echo "MARK start carry" >> gpurun_out/r02b_clocks_c5s.csv
OPKB_CARRY=2 OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C5s-gram > gpurun_out/r02b_c5s_gram_carry.jsonl 2> gpurun_out/r02b_err.log
echo "MARK start sweep" >> gpurun_out/r02b_clocks_c5s.csv
OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C5s-gram > gpurun_out/r02b_c5s_gram.jsonl 2>> gpurun_out/r02b_err.log
echo "MARK start C4" >> gpurun_out/r02b_clocks_c5s.csv
OPKB_BENCH_NATIVE=1 python tools/bench_configs.py C4 > gpurun_out/r02b_c4.jsonl 2>> gpurun_out/r02b_err.log
kill $SMI
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gram2 -s 12 -c 1 -f -o gpurun_out/r02b_gram2_f32 python tools/bench_configs.py C5s-gram > gpurun_out/r02b_ncu1.log 2>&1; echo "ncu1 rc=$?"
python __graft_entry__.py smoke 2>&1 | tail -2
ls -la gpurun_out | grep r02b
