#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t3.log 2>&1
tail -3 gpurun_out/t3.log
python tools/phase_clocks.py ransac_slam_b200/lib/librslam_dbg.so > gpurun_out/phase3.log 2>&1
cat gpurun_out/phase3.log
python bench.py --steps 100 --warmup 5 --extras c5 --no-cpu > gpurun_out/b_c5_3.json 2> gpurun_out/b_c5_3.err
python tools/show_bench.py gpurun_out/b_c5_3.json 2>/dev/null | head -60
