#!/bin/bash
# round-2 GPU session G3 (one B200): Gram sweep, finer role decomposition (OPKB_GRAM_NB) at the C5 shard and at C4
mkdir -p gpurun_out
: > gpurun_out/r02g3_ab.txt
for V in "x=1" "OPKB_GRAM_NB=3" "OPKB_GRAM_NB=4" "OPKB_GRAM_NB=3 OPKB_GRAM_KSHARE=1" "OPKB_GRAM_NB=4 OPKB_GRAM_COOP=1"; do
env $V OPKB_BENCH_NATIVE=1 timeout 200 python tools/bench_configs.py C5s-gram C4 2>gpurun_out/r02g3_err.log | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d.get('kernels_ms_per_iteration',{}); c=d.get('kernel_launches_per_iteration',{})
        print('$V', d['config'][:20], round(d.get('iterations_per_s',0),2), 'gram ms/launch', round(k.get('gram',0)/max(c.get('gram',0),1e-9),3), 'f', d.get('f'), d.get('error'))
" | tee -a gpurun_out/r02g3_ab.txt
tail -2 gpurun_out/r02g3_err.log
done
