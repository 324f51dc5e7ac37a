#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t2.log 2>&1
tail -3 gpurun_out/t2.log
python bench.py --steps 50 --warmup 5 --extras c3 --no-cpu > gpurun_out/b_c3_2.json 2> gpurun_out/b_c3_2.err
python tools/show_bench.py gpurun_out/b_c3_2.json 2>/dev/null | head -60
