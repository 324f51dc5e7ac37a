#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t5.log 2>&1
tail -15 gpurun_out/t5.log
python bench.py --steps 100 --warmup 5 --extras c5 --no-cpu > gpurun_out/b_c5_5.json 2> gpurun_out/b_c5_5.err
python tools/show_bench.py gpurun_out/b_c5_5.json 2>/dev/null | head -60
tail -5 gpurun_out/b_c5_5.err
