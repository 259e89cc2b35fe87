#!/bin/bash
mkdir -p gpurun_out
OPKB_MAILBOX_TIMEOUT=15 timeout 180 python -m pytest tests/test_gpu_group.py -q -x 2>&1 | tail -40 > gpurun_out/r02r_group.log; cat gpurun_out/r02r_group.log | tail -30
