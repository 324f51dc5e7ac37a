#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_map.py -x -q > gpurun_out/t11.log 2>&1
tail -25 gpurun_out/t11.log
python -m pytest tests -m gpu -x -q > gpurun_out/t11b.log 2>&1
tail -5 gpurun_out/t11b.log
