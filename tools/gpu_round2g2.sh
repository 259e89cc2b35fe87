#!/bin/bash
# round-2 GPU session G2 (one B200): Gram sweep with the newest pair formed by all consumer warps -- parity, A/B
mkdir -p gpurun_out
( timeout 800 python -m pytest tests/test_gpu_vmlmb.py tests/test_gpu_carry.py tests/test_gpu_primitives.py -x -q 2>&1 | tail -4 ) > gpurun_out/r02g2_tests.log; tail -2 gpurun_out/r02g2_tests.log
: > gpurun_out/r02g2_ab.txt
for V in "OPKB_GRAM_COOP=0" "OPKB_GRAM_COOP=1"; do
env $V OPKB_BENCH_NATIVE=1 timeout 200 python tools/bench_configs.py C5s-gram C4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['kernels_ms_per_iteration']; c=d['kernel_launches_per_iteration']
        print('$V', d['config'][:20], round(d['iterations_per_s'],2), 'gram ms/launch', round(k['gram']/max(c['gram'],1e-9),3), 'launches/it', c['gram'], 'f', d['f'])
" | tee -a gpurun_out/r02g2_ab.txt
done
