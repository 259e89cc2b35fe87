#!/bin/bash
# round-2 GPU session Y (one B200): tuning knobs of the Float32 Gram sweep at the C5 shard (m = 7)
mkdir -p gpurun_out
: > gpurun_out/r02y_sweep.txt
for V in "x=1" "OPKB_GRAM_KSHARE=3" "OPKB_GRAM_ROWB=3072" "OPKB_GRAM_ROWB=2048" "OPKB_GRAM_KSHARE=3 OPKB_GRAM_ROWB=2048" "OPKB_GRAM_KSHARE=1" "OPKB_GRAM_NSTAGE=2" "OPKB_GRAM_ROWB=2048 OPKB_GRAM_NSTAGE=4"; do
env $V OPKB_BENCH_NATIVE=1 timeout 120 python tools/bench_configs.py C5s-gram 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['kernels_ms_per_iteration']; c=d['kernel_launches_per_iteration']
        print('$V', round(d['iterations_per_s'],1), 'gram ms/launch', round(k['gram']/c['gram'],3), 'dir', round(k['direction']/c['direction'],3), 'trial', round(k['trial']/c['trial'],3))
" | tee -a gpurun_out/r02y_sweep.txt
done
