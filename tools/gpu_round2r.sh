#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_group.py -q -x 2>&1 | tail -40 > gpurun_out/r02r_group.log; cat gpurun_out/r02r_group.log | tail -40
