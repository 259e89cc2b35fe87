#!/bin/bash
# usage: tools/run_multigpu_check.sh N   (under gpurun --gpus N): parity tool, then bench.py with
# the shared-memory mailbox and with the ncclAllGather fall-back; logs under gpurun_out/
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
# second argument "bench": skip the parity tool and the e2e leg (large N is charged N times);
# "parity": the parity tool only
E2E=""
if [ "$2" != "bench" ]; then
$TR --master-port 29511 tools/multigpu_parity.py > gpurun_out/parity_n$N.log 2>&1; echo rc=$? >> gpurun_out/parity_n$N.log
else
E2E="--no-e2e"; : > gpurun_out/parity_n$N.log
fi
if [ "$2" = "parity" ]; then grep "GPUs\]\|rc=\|Error\|error" gpurun_out/parity_n$N.log | tail -20; exit 0; fi
$TR --master-port 29513 bench.py --gpus $N $E2E > gpurun_out/bench_n${N}_mbox.json 2> gpurun_out/bench_n${N}_mbox.err
OPKB_MAILBOX=0 $TR --master-port 29514 bench.py --gpus $N --no-e2e > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err
grep "GPUs\]\|rc=\|Error\|error" gpurun_out/parity_n$N.log | tail -20
for f in gpurun_out/bench_n${N}_mbox.json gpurun_out/bench_n${N}_nccl.json; do
python - "$f" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print(sys.argv[1], "it/s", round(d["value"], 2), "ms/it", round(d["ms_per_step"], 4), "kernel ms",
              round(d["roofline"]["kernels"]["direction"]["avg_ms"], 4), "e2e", (d.get("e2e") or {}).get("value"),
              d["config"]["sharding"][-70:])
PY
done
tail -3 gpurun_out/bench_n${N}_mbox.err
